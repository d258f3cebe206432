#!/bin/bash
# round 2, visit p: level-1 hidden column by cp.async.bulk + mbarrier (A/B against the producers' cp.async) + full suite
set -u
OUT=gpurun_out/r02p; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
BENCH_STEPS=10 tools/gpu_try.sh r02p "" "bulk:--torch-gpu-rotations 0" "nobulk:lib=build/variants/libdlpd_nobulk.so --torch-gpu-rotations 0" "bulk2:--torch-gpu-rotations 0"
