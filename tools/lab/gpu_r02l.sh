#!/bin/bash
# round 2, visit l: timing ablations of the tcgen05 epilogue (results are wrong by construction, only the clock is read)
set -u
BENCH_STEPS=6 tools/gpu_try.sh r02l "" "default:--torch-gpu-rotations 0 --rot-batch 64" "abl_load:lib=build/variants/libdlpd_abl_load.so --torch-gpu-rotations 0 --rot-batch 64" \
  "abl_fft:lib=build/variants/libdlpd_abl_fft.so --torch-gpu-rotations 0 --rot-batch 64" "abl_mlp:lib=build/variants/libdlpd_abl_mlp.so --torch-gpu-rotations 0 --rot-batch 64" \
  "abl_fftmlp:lib=build/variants/libdlpd_abl_fftmlp.so --torch-gpu-rotations 0 --rot-batch 64"
