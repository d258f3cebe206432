#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list and full captures.  Everything lands in gpurun_out/.
# usage: tools/gpu_round.sh [tag] [stages...]   stages default: tests smoke bench launches ncu
set -u
TAG=${1:-r01}; shift || true
STAGES=${*:-tests smoke bench launches ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > $OUT/gpu.csv 2>&1
python -c "import os; print('cpu_count', os.cpu_count())" >> $OUT/gpu.csv
for s in $STAGES; do
  case $s in
    tests)
      timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log ;;
    tests_all)
      timeout 1200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log ;;
    bench)
      timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/bench.log 2>&1; echo "bench rc=$?" >> $OUT/bench.log ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dlpd_kernel -c 60 --csv \
        --log-file $OUT/launches.csv python bench.py --steps 1 --warmup 1 --cpu-rotations 0 --rot-batch 16 --max-rot 16 \
        > $OUT/launches_run.log 2>&1 ;;
    ubench)
      for b in fp32x2_rate mma_tf32_rate; do ./build/$b; done > $OUT/ubench.log 2>&1 ;;
    ncu)
      timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
        -k "regex:Epilogue|Plane|FwdZC|HiddenL1" -s 2 -c 6 -o $OUT/prof -f \
        python bench.py --steps 1 --warmup 0 --cpu-rotations 0 --rot-batch 32 --max-rot 32 > $OUT/ncu_run.log 2>&1 ;;
    sanitize)
      timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x \
        -k "(small and 8-4) or ties or (cluster_plane_search_small and 16-4) or (pipelined_epilogue_small and 8-4) or project_clash or (tensor_filter_small and 8-16) or volume_convolution_128 or test_search_full_size_80_matches" \
        > $OUT/memcheck.log 2>&1
      timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -x \
        -k "(search_small_matches_oracle and 8-4) or (cluster_plane_search_small and 16-4) or (pipelined_epilogue_small and 8-4) or (tensor_filter_small and 8-16) or volume_convolution_128 or test_search_full_size_80_matches" \
        > $OUT/racecheck.log 2>&1 ;;
    cfg5)
      timeout 600 python bench.py --config 5 --steps 4 --warmup 3 --cpu-rotations 0 --rot-batch 32 --max-rot 32 > $OUT/bench_cfg5.log 2>&1; echo "rc=$?" >> $OUT/bench_cfg5.log ;;
    cfg4)
      timeout 600 python bench.py --config 4 --steps 8 --warmup 3 --cpu-rotations 0 > $OUT/bench_cfg4.log 2>&1; echo "rc=$?" >> $OUT/bench_cfg4.log ;;
    cfg3)
      timeout 600 python bench.py --config 3 --steps 8 --warmup 3 --cpu-rotations 0 > $OUT/bench_cfg3.log 2>&1; echo "rc=$?" >> $OUT/bench_cfg3.log ;;
    scale2)
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_2gpu.log 2>&1 ;;
  esac
done
tail -n 5 $OUT/*.log
