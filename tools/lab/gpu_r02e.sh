#!/bin/bash
# round 2, visit e: TMA-fed tcgen05 epilogue: full GPU suite + bench A/B
set -u
OUT=gpurun_out/r02e; mkdir -p $OUT
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log; tail -3 $OUT/smoke.log
timeout 1800 python -m pytest tests -m gpu -q --durations=10 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -n 40 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 --torch-gpu-rotations 0 > $OUT/bench.log 2>&1; echo "bench rc=$?" >> $OUT/bench.log
python - <<PY
import json
for name in ("bench",):
    try:
        l=[x for x in open("$OUT/%s.log" % name) if x.startswith("{")][-1]; d=json.loads(l)
        print(name, "rot/s %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: round(v["ms_per_launch"],3) for k,v in d["kernels"].items()})
        print("parity", d.get("parity")); print("best", d.get("best_pose")); print("clocks", d.get("clocks")); print("fp32", d.get("roofline_fp32"))
    except Exception as e:
        print("failed", e); print(open("$OUT/%s.log" % name).read()[-1500:])
PY
