#!/bin/bash
# Build variants of libdlpd.so ON THE GPU BOX with different -D tunables and bench each (short runs).
# usage: [BENCH_ARGS="--config 5"] tools/gpu_variants.sh tag "name1:-DX=1 -DY=2" "name2:..."
set -u
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  so=/tmp/libdlpd_$name.so
  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared $flags \
     -o $so deeplocalproteindocking_b200/csrc/dlpd_api.cu > $OUT/build_$name.log 2>&1 || { echo "build failed $name"; continue; }
  DLPD_LIB_PATH=$so timeout 300 python bench.py --steps ${BENCH_STEPS:-6} --warmup 3 --cpu-rotations 0 ${BENCH_ARGS:-} > $OUT/bench_$name.log 2>&1
  echo "== $name ($flags)"; python - <<PY
import json
try:
    l=[x for x in open("$OUT/bench_$name.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("rot/s %.1f  e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: round(v["ms_per_launch"],3) for k,v in d["kernels"].items()})
except Exception as e:
    print("failed", e); print(open("$OUT/bench_$name.log").read()[-800:])
PY
done
