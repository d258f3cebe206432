#!/bin/bash
# round 2, visit q (8 GPUs): strong scaling of config 2 at N = 8 / 4, config 3 (oim04 surrogate) full set on 8 GPUs, configs 4 and 5
# on 8 GPUs, two-rank NCCL identity test
set -u
OUT=gpurun_out/r02q; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpu.csv
run() { name=$1; n=$2; shift 2; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $n "$@" > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $OUT/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_multi.log; tail -2 $OUT/pytest_multi.log
run bench_cfg2_8gpu 8 --steps 20 --warmup 3
run bench_cfg2_4gpu 4 --steps 20 --warmup 3
run bench_cfg3_8gpu_full 8 --config 3 --full-set --steps 20 --warmup 2
run bench_cfg4_8gpu 8 --config 4 --steps 8 --warmup 3
run bench_cfg5_8gpu 8 --config 5 --steps 4 --warmup 3 --rot-batch 32
python - <<PY
import json
for name in ("bench_cfg2_8gpu", "bench_cfg2_4gpu", "bench_cfg3_8gpu_full", "bench_cfg4_8gpu", "bench_cfg5_8gpu"):
    try:
        l=[x for x in open("$OUT/%s.log" % name) if x.startswith("{")][-1]; d=json.loads(l)
        print(name, "rot/s %.1f e2e %.1f total_ms %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]*d["steps"]), d["scaling"], d["step"]["rotations_total"], d["clocks"], d.get("best_pose"))
    except Exception as e:
        print("failed", name, e); print(open("$OUT/%s.log" % name).read()[-1200:])
PY
