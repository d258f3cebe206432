#!/bin/bash
# round 2, visit zz (8 GPUs): the driver's scaling sequence on the final build -- config 2 at N = 1, 2, 4, 8 back to back
# (steps 20, warm-up 5), reference arm at N = 8 (rank 0 only works), two-rank NCCL identity test
set -u
OUT=gpurun_out/r02zz; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/gpu.csv
run() { name=$1; n=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $n "$@" > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $OUT/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_multi.log; tail -2 $OUT/pytest_multi.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/scale_1gpu.log 2>&1; echo "rc=$?" >> $OUT/scale_1gpu.log
run scale_2gpu 2 --steps 20 --warmup 5
run scale_4gpu 4 --steps 20 --warmup 5
run scale_8gpu 8 --steps 20 --warmup 5
run reference_8gpu 8 --impl reference --steps 3 --warmup 1
python - <<PY
import json
for name in ("scale_1gpu", "scale_2gpu", "scale_4gpu", "scale_8gpu", "reference_8gpu"):
    try:
        l=[x for x in open("$OUT/%s.log" % name) if x.startswith("{")][-1]; d=json.loads(l)
        print(name, "rot/s %.1f e2e %.1f total_ms %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]*d["steps"]), d.get("scaling"), d.get("clocks"), d.get("best_pose"))
    except Exception as e:
        print("failed", name, e); print(open("$OUT/%s.log" % name).read()[-1200:])
PY
