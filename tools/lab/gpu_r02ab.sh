#!/bin/bash
# round 2, visit ab: Plane<64> in 4 teams + HiddenL1 2-line tiles at 128^3 as defaults -- GPU suite, config 5 bench + parity sweep,
# default bench
set -u
OUT=gpurun_out/r02ab; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py --config 5 --steps 4 --warmup 2 --rot-batch 32 > $OUT/bench_cfg5.log 2>&1; echo "rc=$?" >> $OUT/bench_cfg5.log
timeout 900 python tests/sweeps/parity_sweep.py --config 5 --rotations 32 --seed 77 > $OUT/parity_sweep_cfg5.log 2>&1; echo "rc=$?" >> $OUT/parity_sweep_cfg5.log; tail -2 $OUT/parity_sweep_cfg5.log | cut -c1-400
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench.log 2>&1; echo "bench rc=$?" >> $OUT/bench.log
python - <<PY
import json
for name in ("bench_cfg5", "bench"):
    try:
        l=[x for x in open("$OUT/%s.log" % name) if x.startswith("{")][-1]; d=json.loads(l)
        print(name, "value %.2f e2e %.2f ms/step %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("parity"), "  ".join("%s %.2f" % (n, v["ms_per_launch"]) for n, v in d["kernels"].items()))
    except Exception as e:
        print("failed", name, e); print(open("$OUT/%s.log" % name).read()[-1500:])
PY
