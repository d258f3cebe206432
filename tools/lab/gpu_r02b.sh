#!/bin/bash
# round 2, visit b: tcgen05 filter microbenchmark variants and ablations
set -u
OUT=gpurun_out/r02b; mkdir -p $OUT
for m in 0 1 2 3 11 12 14 17 21; do
  timeout 60 ./build/tc_filter $m 50 > $OUT/tc_filter_$m.log 2>&1; echo "rc=$?" >> $OUT/tc_filter_$m.log
done
timeout 60 ./build/tc_filter 3 0.01 > $OUT/tc_filter_3_small.log 2>&1
timeout 60 ./build/tc_filter 3 30000 > $OUT/tc_filter_3_big.log 2>&1
cat $OUT/tc_filter_*.log
