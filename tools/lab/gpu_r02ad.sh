#!/bin/bash
# round 2, visit ad: compute-sanitizer on the final build -- memcheck / racecheck / synccheck over the 128^3 search (new Plane<64>
# teams, HiddenL1 2-line tiles, 1024-thread FwdZC), the 80^3 search on the tcgen05 epilogue and the atoms path
set -u
OUT=gpurun_out/r02ad; mkdir -p $OUT
K='test_search_full_size_128_matches or test_search_full_size_80_matches or (search_small_matches_oracle and 8-4)'
K2='bench_inputs or typed'
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "$K" > $OUT/memcheck.log 2>&1; echo "rc=$?" >> $OUT/memcheck.log; tail -4 $OUT/memcheck.log
timeout 1200 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x -k "$K" > $OUT/racecheck.log 2>&1; echo "rc=$?" >> $OUT/racecheck.log; tail -4 $OUT/racecheck.log
timeout 1200 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -q -x -k "$K" > $OUT/synccheck.log 2>&1; echo "rc=$?" >> $OUT/synccheck.log; tail -4 $OUT/synccheck.log
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_round2.py -q -x -k "$K2" > $OUT/memcheck_round2.log 2>&1; echo "rc=$?" >> $OUT/memcheck_round2.log; tail -4 $OUT/memcheck_round2.log
