#!/bin/bash
# round 2, visit v: FwdZC with 16-channel records (64 bytes per tap) and 4-line tiles
set -u
OUT=gpurun_out/r02v; mkdir -p $OUT
DLPD_LIB_PATH=$PWD/build/variants/libdlpd_cg16a.so timeout 600 python -m pytest tests -m gpu -q -x -k "full_size_80_matches or (small_matches and 20-16)" > $OUT/pytest_cg16.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_cg16.log; tail -3 $OUT/pytest_cg16.log
BENCH_STEPS=10 tools/gpu_try.sh r02v "" "cg16a:lib=build/variants/libdlpd_cg16a.so --torch-gpu-rotations 0" "cg16b:lib=build/variants/libdlpd_cg16b.so --torch-gpu-rotations 0"
