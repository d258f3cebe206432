#!/bin/bash
# round 2, visit t: pipelined epilogue with the tcgen05 scoring stage (experimental build): parity + A/B bench
set -u
OUT=gpurun_out/r02t; mkdir -p $OUT
export DLPD_LIB_PATH=$PWD/build/variants/libdlpd_exp.so
timeout 900 python -m pytest tests -m gpu -q -x -k "pipe_tc" > $OUT/pytest_pipe_tc.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_pipe_tc.log; tail -5 $OUT/pytest_pipe_tc.log
unset DLPD_LIB_PATH
BENCH_STEPS=10 tools/gpu_try.sh r02t "" "default:--torch-gpu-rotations 0" "pipe_tc:lib=build/variants/libdlpd_exp.so --epilogue-kind pipe_tc --torch-gpu-rotations 0"
