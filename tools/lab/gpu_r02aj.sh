#!/bin/bash
# round 2, visit aj (8 GPUs): configs 5 and 4 on 8 GPUs with the final build
set -u
OUT=gpurun_out/r02aj; mkdir -p $OUT
run() { name=$1; n=$2; shift 2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $n "$@" > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log; }
run bench_cfg5_8gpu 8 --config 5 --steps 4 --warmup 3 --rot-batch 32 --cpu-rotations 0 --torch-gpu-rotations 0
run bench_cfg4_8gpu 8 --config 4 --steps 8 --warmup 3 --cpu-rotations 0 --torch-gpu-rotations 0
python - <<PY
import json
for name in ("bench_cfg5_8gpu", "bench_cfg4_8gpu"):
    try:
        l=[x for x in open("$OUT/%s.log" % name) if x.startswith("{")][-1]; d=json.loads(l)
        print(name, "rot/s %.1f e2e %.1f total_ms %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]*d["steps"]), d["scaling"], d["step"]["rotations_total"], d["clocks"])
    except Exception as e:
        print("failed", name, e); print(open("$OUT/%s.log" % name).read()[-1200:])
PY
