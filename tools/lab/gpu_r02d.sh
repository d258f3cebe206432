#!/bin/bash
# round 2, visit d: new round-2 GPU tests, docker tests, smoke, new bench line (full oim10 set)
set -u
OUT=gpurun_out/r02d; mkdir -p $OUT
python -c "import os; print('cpu_count', os.cpu_count())" > $OUT/gpu.csv
timeout 1500 python -m pytest tests/test_gpu_round2.py tests/test_gpu_docker.py -m gpu -q --durations=15 > $OUT/pytest_round2.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_round2.log
tail -n 45 $OUT/pytest_round2.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log; tail -3 $OUT/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.log 2>&1; echo "bench rc=$?" >> $OUT/bench.log
tail -c 6000 $OUT/bench.log
