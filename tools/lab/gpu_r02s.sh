#!/bin/bash
# round 2, visit s: wide model (32 + 64 channels, H = 48), full suite, bench
set -u
OUT=gpurun_out/r02s; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -16 $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.log 2>&1; echo "bench rc=$?" >> $OUT/bench.log
python - <<PY
import json
l=[x for x in open("$OUT/bench.log") if x.startswith("{")][-1]; d=json.loads(l)
print("rot/s %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: round(v["ms_per_launch"],3) for k,v in d["kernels"].items()})
print(d["parity"]); print(d["step"]); print(d["clocks"]); print(d["roofline"]["frac"], d["roofline_step"]["frac"], d["roofline_fp32"]["frac"]); print(d["cpu_baseline"]); print(d.get("gpu_torch_baseline"))
PY
