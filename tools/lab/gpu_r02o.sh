#!/bin/bash
# round 2, visit o: compute-sanitizer on the new kernels (tcgen05 epilogue, typed projector, atoms path)
set -u
OUT=gpurun_out/r02o; mkdir -p $OUT
K="(filter_kind_small and 8-16) or (filter_kind_small and 20-16) or typed_projector or plan_options or (adversarial and delta)"
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x -k "$K" > $OUT/memcheck.log 2>&1; echo "rc=$?" >> $OUT/memcheck.log
timeout 1200 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -x -k "(filter_kind_small and 8-16) or (filter_kind_small and 20-16)" > $OUT/racecheck.log 2>&1; echo "rc=$?" >> $OUT/racecheck.log
timeout 1200 compute-sanitizer --tool synccheck python -m pytest tests -m gpu -q -x -k "(filter_kind_small and 8-16) or (filter_kind_small and 20-16)" > $OUT/synccheck.log 2>&1; echo "rc=$?" >> $OUT/synccheck.log
tail -n 6 $OUT/memcheck.log $OUT/racecheck.log $OUT/synccheck.log
