#!/bin/bash
# round 2, visit x: Plane<80> without the two team barriers of its y phase (timing ablation, wrong results)
set -u
BENCH_STEPS=6 tools/gpu_try.sh r02x "" "default:--torch-gpu-rotations 0 --rot-batch 64" "abl_ybar:lib=build/variants/libdlpd_abl_ybar.so --torch-gpu-rotations 0 --rot-batch 64"
