#!/bin/bash
set -u
BENCH_STEPS=6 tools/gpu_try.sh r02m "" "abl_all:lib=build/variants/libdlpd_abl_all.so --torch-gpu-rotations 0 --rot-batch 64" \
  "abl_loadfft:lib=build/variants/libdlpd_abl_loadfft.so --torch-gpu-rotations 0 --rot-batch 64" "abl_loadmlp:lib=build/variants/libdlpd_abl_loadmlp.so --torch-gpu-rotations 0 --rot-batch 64"
