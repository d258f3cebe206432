#!/bin/bash
# round 2, visit a: tcgen05 filter microbenchmark (numerics + time per tile) and the prepared packed-twiddle builds
set -u
OUT=gpurun_out/r02a; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > $OUT/gpu.csv 2>&1
for m in 0 1 2; do
  timeout 60 ./build/tc_filter $m 50 > $OUT/tc_filter_$m.log 2>&1; echo "rc=$?" >> $OUT/tc_filter_$m.log
done
timeout 60 ./build/tc_filter 1 0.01 > $OUT/tc_filter_1_small.log 2>&1; echo "rc=$?" >> $OUT/tc_filter_1_small.log
timeout 60 ./build/tc_filter 1 30000 > $OUT/tc_filter_1_big.log 2>&1; echo "rc=$?" >> $OUT/tc_filter_1_big.log
cat $OUT/tc_filter_*.log
BENCH_STEPS=6 tools/gpu_try.sh r02a "" "base:" "pkepi:lib=build/variants/libdlpd_pkepi.so" "pkfwd:lib=build/variants/libdlpd_pkfwd.so" \
  "pkhid:lib=build/variants/libdlpd_pkhid.so" "pkall:lib=build/variants/libdlpd_pkall.so"
