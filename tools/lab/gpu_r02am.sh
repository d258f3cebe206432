#!/bin/bash
# round 2, visit am: config 3 (oim04 surrogate, 232 020 rotations) in full on ONE GPU -- a 90-second sustained run
set -u
OUT=gpurun_out/r02am; mkdir -p $OUT
timeout 900 python bench.py --config 3 --full-set --steps 20 --warmup 2 --cpu-rotations 0 --torch-gpu-rotations 0 > $OUT/bench_cfg3_full_1gpu.log 2>&1; echo "rc=$?" >> $OUT/bench_cfg3_full_1gpu.log
python - <<PY
import json
l=[x for x in open("$OUT/bench_cfg3_full_1gpu.log") if x.startswith("{")][-1]; d=json.loads(l)
print("rot/s %.1f e2e %.1f total_s %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]*d["steps"]/1e3), d["step"]["rotations_total"], d["clocks"])
PY
