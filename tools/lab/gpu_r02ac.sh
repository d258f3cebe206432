#!/bin/bash
# round 2, visit ac: config 5 (128^3) with FwdZC tile / thread variants (libraries under variants/, DLPD_LIB_PATH)
set -u
OUT=gpurun_out/r02ac; mkdir -p $OUT
one() { name=$1; lib=$2; DLPD_LIB_PATH=$lib timeout 600 python bench.py --config 5 --steps 4 --warmup 2 --rot-batch 32 --cpu-rotations 0 --torch-gpu-rotations 0 > $OUT/cfg5_$name.log 2>&1; echo "rc=$?" >> $OUT/cfg5_$name.log; }
one default deeplocalproteindocking_b200/libdlpd.so
for v in fa fb fc fd fe; do one $v variants/libdlpd_$v.so; done
python - <<PY
import json
for name in "default fa fb fc fd fe".split():
    try:
        l=[x for x in open("$OUT/cfg5_%s.log" % name) if x.startswith("{")][-1]; d=json.loads(l)
        k=d["kernels"]
        print("%-8s rot/s %.1f  " % (name, d["value"]) + "  ".join("%s %.2f" % (n.replace("_level","L"), v["ms_per_launch"]) for n, v in k.items()))
    except Exception as e:
        print("failed", name, e); print(open("$OUT/cfg5_%s.log" % name).read()[-800:])
PY
