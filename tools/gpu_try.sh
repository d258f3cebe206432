#!/bin/bash
# Quick A/B visit: GPU parity tests (optionally filtered), then short bench runs with different runtime switches / builds.
# usage: tools/gpu_try.sh tag "pytest -k expr | all | empty" "name:[lib=path.so] bench args" ...
set -u
TAG=$1; KEXPR=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
if [ -n "$KEXPR" ]; then
  if [ "$KEXPR" = all ]; then timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
  else timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" > $OUT/pytest_gpu.log 2>&1; fi
  echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -n 15 $OUT/pytest_gpu.log
fi
for b in ${UBENCH:-}; do ./build/$b > $OUT/ubench_$b.log 2>&1; cat $OUT/ubench_$b.log; done
for v in "$@"; do
  name=${v%%:*}; args=${v#*:}; lib=""
  case "$args" in lib=*) lib=${args%% *}; lib=${lib#lib=}; args=${args#* }; [ "$args" = "lib=$lib" ] && args="";; esac
  DLPD_LIB_PATH=${lib:+$PWD/$lib} timeout 300 python bench.py --steps ${BENCH_STEPS:-6} --warmup 3 --cpu-rotations 0 $args > $OUT/bench_$name.log 2>&1
  echo "== $name (${lib:-default build} $args)"; python - <<PY
import json
try:
    l=[x for x in open("$OUT/bench_$name.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("rot/s %.1f  e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: round(v["ms_per_launch"],3) for k,v in d["kernels"].items()})
except Exception as e:
    print("failed", e); print(open("$OUT/bench_$name.log").read()[-800:])
PY
done
