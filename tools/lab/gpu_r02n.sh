#!/bin/bash
# round 2, visit n: z lines of the 80^3 box factored 8 x 20 instead of 10 x 16
set -u
OUT=gpurun_out/r02n; mkdir -p $OUT
DLPD_LIB_PATH=$PWD/build/variants/libdlpd_alt160.so timeout 900 python -m pytest tests -m gpu -q -x -k "full_size_80 or filter_kind_80 or bench_inputs or pair_switch" > $OUT/pytest_alt160.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_alt160.log; tail -4 $OUT/pytest_alt160.log
BENCH_STEPS=10 tools/gpu_try.sh r02n "" "default:--torch-gpu-rotations 0" "alt160:lib=build/variants/libdlpd_alt160.so --torch-gpu-rotations 0"
