#!/bin/bash
# round 2, visit r: HiddenL1's 32 -> 24 contraction on tcgen05: parity (full suite), A/B bench, 512-rotation parity sweep
set -u
OUT=gpurun_out/r02r; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
BENCH_STEPS=20 tools/gpu_try.sh r02r "" "hidtc:--torch-gpu-rotations 0" "nohidtc:lib=build/variants/libdlpd_nohidtc.so --torch-gpu-rotations 0"
timeout 1200 python tests/sweeps/parity_sweep.py --rotations 512 > $OUT/parity_sweep.log 2>&1; echo "rc=$?" >> $OUT/parity_sweep.log; tail -c 1200 $OUT/parity_sweep.log
