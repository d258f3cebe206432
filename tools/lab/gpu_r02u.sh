#!/bin/bash
set -u
BENCH_STEPS=10 tools/gpu_try.sh r02u "" "pipe416:lib=build/variants/libdlpd_exp416.so --epilogue-kind pipe_tc --torch-gpu-rotations 0" "pipe448:lib=build/variants/libdlpd_exp448.so --epilogue-kind pipe_tc --torch-gpu-rotations 0"
