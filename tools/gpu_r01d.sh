#!/bin/bash
# cluster plane + 128^3: parity tests, then bench single vs cluster plane at L=80
set -u
OUT=gpurun_out/r01d; mkdir -p $OUT
free -g > $OUT/host.txt; nproc >> $OUT/host.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
for mode in single cluster; do
  timeout 300 python bench.py --steps 6 --warmup 3 --cpu-rotations 0 --plane-mode $mode > $OUT/bench_$mode.log 2>&1
done
mkdir -p build; for b in fp32x2_rate mma_tf32_rate; do nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/$b tools/ubench/$b.cu && ./build/$b; done > $OUT/ubench.log 2>&1
tail -n 4 $OUT/*.log
