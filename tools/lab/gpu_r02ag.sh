#!/bin/bash
# round 2, visit ag: Docker.dockSE3 with the rotation matrices uploaded once (no pageable copy per batch): docker tests, top-K run
# at K = 2000 with 32 and 64 rotations per fused call
set -u
OUT=gpurun_out/r02ag; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_docker.py tests/test_gpu_round2.py -m gpu -q -x -k "dock" > $OUT/pytest_docker.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_docker.log; tail -3 $OUT/pytest_docker.log
timeout 1200 python tests/sweeps/topk_bench.py --rotations 4096 --k 2000 --parity-rotations 96 > $OUT/topk_rb32.log 2>&1; echo "topk rc=$?" >> $OUT/topk_rb32.log
tail -c 1800 $OUT/topk_rb32.log
timeout 1200 python tests/sweeps/topk_bench.py --rotations 4096 --k 2000 --parity-rotations 0 --rot-batch 64 > $OUT/topk_rb64.log 2>&1; echo "topk rc=$?" >> $OUT/topk_rb64.log
tail -c 1200 $OUT/topk_rb64.log
