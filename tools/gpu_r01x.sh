#!/bin/bash
set -u
OUT=gpurun_out/r01x; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "duo or kinds" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
for k in classic duo; do
  timeout 300 python bench.py --steps 6 --warmup 3 --cpu-rotations 0 --epilogue-kind $k > $OUT/bench_$k.log 2>&1
done
tail -n 3 $OUT/pytest_gpu.log
python - <<PY
import json
for f in ("bench_classic","bench_duo"):
    try:
        l=[x for x in open("$OUT/%s.log" % f) if x.startswith("{")][-1]; d=json.loads(l)
        print(f, round(d["value"],1), {k: round(v["ms_per_launch"],3) for k,v in d["kernels"].items()})
    except Exception as e: print(f, "failed", e)
PY
