#!/bin/bash
# round 2, visit c: microbenchmark variants, full GPU test-suite with the tcgen05 epilogue, A/B bench of the filter kinds
set -u
OUT=gpurun_out/r02c; mkdir -p $OUT
for m in 0 1 2 3 11 12 14 17 21; do
  timeout 60 ./build/tc_filter $m 50 > $OUT/tc_filter_$m.log 2>&1; echo "rc=$?" >> $OUT/tc_filter_$m.log
done
timeout 60 ./build/tc_filter 3 0.01 > $OUT/tc_filter_3_small.log 2>&1
timeout 60 ./build/tc_filter 3 30000 > $OUT/tc_filter_3_big.log 2>&1
cat $OUT/tc_filter_*.log | grep -v "^rc=0"
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -n 25 $OUT/pytest_gpu.log
BENCH_STEPS=6 tools/gpu_try.sh r02c "" "fp32:--filter-kind fp32" "tcgen05:--filter-kind tcgen05"
