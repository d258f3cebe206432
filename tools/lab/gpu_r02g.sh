#!/bin/bash
# round 2, visit g (2 GPUs): strong-scaling bench at N = 2 and the two-rank NCCL identity test; reference arm
set -u
OUT=gpurun_out/r02g; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpu.csv
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $OUT/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_multi.log; tail -3 $OUT/pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/bench_2gpu.log 2>&1; echo "rc=$?" >> $OUT/bench_2gpu.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --cpu-rotations 0 --torch-gpu-rotations 0 > $OUT/bench_1gpu.log 2>&1; echo "rc=$?" >> $OUT/bench_1gpu.log
python - <<PY
import json
for name in ("bench_1gpu", "bench_2gpu"):
    try:
        l=[x for x in open("$OUT/%s.log" % name) if x.startswith("{")][-1]; d=json.loads(l)
        print(name, "rot/s %.1f e2e %.1f ms_total %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]*d["steps"]), d["scaling"], d["step"], d["setup_ms"], d["clocks"], d.get("best_pose"))
    except Exception as e:
        print("failed", name, e); print(open("$OUT/%s.log" % name).read()[-1500:])
PY
