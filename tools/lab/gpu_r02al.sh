#!/bin/bash
# round 2, visit ae/af: config 5 A/B of cluster-barrier variants of PlaneX (variants/*.so) vs the in-tree build,
# then the 128^3 parity tests on the variant
set -u
OUT=gpurun_out/r02al; mkdir -p $OUT
one() { name=$1; lib=$2; DLPD_LIB_PATH=$lib timeout 600 python bench.py --config 5 --steps 4 --warmup 2 --rot-batch 32 --cpu-rotations 0 --torch-gpu-rotations 0 > $OUT/cfg5_$name.log 2>&1; echo "rc=$?" >> $OUT/cfg5_$name.log; }
one default deeplocalproteindocking_b200/libdlpd.so
one nolast variants/libdlpd_nolast.so
one default2 deeplocalproteindocking_b200/libdlpd.so
one nolast2 variants/libdlpd_nolast.so
for l in deeplocalproteindocking_b200/libdlpd.so variants/libdlpd_nolast.so; do DLPD_LIB_PATH=$l timeout 900 python -m pytest tests -m gpu -q -x -k "128 or cluster" > $OUT/pytest_$(basename $l .so).log 2>&1; echo "rc=$?" >> $OUT/pytest_$(basename $l .so).log; tail -2 $OUT/pytest_$(basename $l .so).log; done
python - <<PY
import json
for name in "default nolast default2 nolast2".split():
    try:
        l=[x for x in open("$OUT/cfg5_%s.log" % name) if x.startswith("{")][-1]; d=json.loads(l)
        k=d["kernels"]
        print("%-8s rot/s %.1f  " % (name, d["value"]) + "  ".join("%s %.2f" % (n.replace("_level","L"), v["ms_per_launch"]) for n, v in k.items()))
    except Exception as e:
        print("failed", name, e); print(open("$OUT/cfg5_%s.log" % name).read()[-800:])
PY
