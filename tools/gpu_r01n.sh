#!/bin/bash
set -u
OUT=gpurun_out/r01n; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "pipelined" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
for k in classic pipelined; do
  timeout 300 python bench.py --steps 6 --warmup 3 --cpu-rotations 0 --epilogue-kind $k > $OUT/bench_$k.log 2>&1
done
tail -n 3 $OUT/*.log | cut -c1-300
