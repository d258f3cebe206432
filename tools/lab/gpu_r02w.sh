#!/bin/bash
# round 2, visit w: what the driver runs at round end -- GPU suite, smoke, default bench (both arms)
set -u
OUT=gpurun_out/r02w; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log; tail -2 $OUT/smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $OUT/bench_reference.log 2>&1; echo "rc=$?" >> $OUT/bench_reference.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > $OUT/bench.log 2>&1; echo "bench rc=$?" >> $OUT/bench.log
timeout 900 python bench.py > $OUT/bench_noflags.log 2>&1; echo "bench rc=$?" >> $OUT/bench_noflags.log
python - <<PY
import json
for name in ("bench_reference", "bench", "bench_noflags"):
    try:
        l=[x for x in open("$OUT/%s.log" % name) if x.startswith("{")][-1]; d=json.loads(l)
        print(name, "value %.2f e2e %.2f steps %d ms/step %.1f" % (d["value"], d["e2e"]["value"], d["steps"], d["ms_per_step"]), d.get("parity"), d.get("clocks"))
    except Exception as e:
        print("failed", name, e); print(open("$OUT/%s.log" % name).read()[-1500:])
PY
