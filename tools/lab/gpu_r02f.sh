#!/bin/bash
# round 2, visit f: tcgen05 epilogue with transform helper warps (cp.async producers), top-K at K = 2000
set -u
OUT=gpurun_out/r02f; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x -k "filter_kind or small_matches or docker or adversarial or plan_options or search_128 or typed" > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -n 8 $OUT/pytest_gpu.log
BENCH_STEPS=10 tools/gpu_try.sh r02f "" "nx64:--torch-gpu-rotations 0" "nx0:lib=build/variants/libdlpd_nx0.so --torch-gpu-rotations 0" "fp32:--filter-kind fp32 --torch-gpu-rotations 0"
timeout 1200 python tests/sweeps/topk_bench.py --rotations 2048 --k 2000 --parity-rotations 64 > $OUT/topk.log 2>&1; echo "topk rc=$?" >> $OUT/topk.log
tail -c 3000 $OUT/topk.log
