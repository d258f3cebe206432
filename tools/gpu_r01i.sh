#!/bin/bash
set -u
OUT=gpurun_out/r01i; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q -k "cluster or 128" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 300 python bench.py --steps 6 --warmup 3 --cpu-rotations 0 --plane-mode cluster > $OUT/bench_cluster.log 2>&1
timeout 600 python bench.py --config 5 --steps 4 --warmup 3 --cpu-rotations 0 --rot-batch 32 --max-rot 32 > $OUT/bench_cfg5.log 2>&1
tail -n 3 $OUT/*.log | cut -c1-300
