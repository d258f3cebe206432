#!/bin/bash
# round 2, visit k: producer-thread counts of the tcgen05 epilogue
set -u
BENCH_STEPS=10 tools/gpu_try.sh r02k "" "default:--torch-gpu-rotations 0" "np96:lib=build/variants/libdlpd_np96.so --torch-gpu-rotations 0" \
  "np128:lib=build/variants/libdlpd_np128.so --torch-gpu-rotations 0" "np96x0:lib=build/variants/libdlpd_np96x0.so --torch-gpu-rotations 0" \
  "np64x32:lib=build/variants/libdlpd_np64x32.so --torch-gpu-rotations 0"
