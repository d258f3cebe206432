#!/bin/bash
# round 2, visit y: asynchronous candidate batches in Docker.dockSE3: docker tests + top-K run at K = 2000
set -u
OUT=gpurun_out/r02y; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_docker.py tests/test_gpu_round2.py -m gpu -q -x -k "dock" > $OUT/pytest_docker.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_docker.log; tail -3 $OUT/pytest_docker.log
timeout 1200 python tests/sweeps/topk_bench.py --rotations 4096 --k 2000 --parity-rotations 96 > $OUT/topk.log 2>&1; echo "topk rc=$?" >> $OUT/topk.log
tail -c 2500 $OUT/topk.log
