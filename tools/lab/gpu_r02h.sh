#!/bin/bash
# round 2, visit h: two tcgen05 half-tile epilogue CTAs per SM (tc_duo): parity + A/B bench
set -u
OUT=gpurun_out/r02h; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "tc_duo or filter_kind" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -n 8 $OUT/pytest_gpu.log
BENCH_STEPS=10 tools/gpu_try.sh r02h "" "default:--torch-gpu-rotations 0" "tcduo:--epilogue-kind tc_duo --torch-gpu-rotations 0"
