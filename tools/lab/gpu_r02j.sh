#!/bin/bash
# round 2, visit j: full GPU suite, epilogue thread-shape variants, configs 4 and 5
set -u
OUT=gpurun_out/r02j; mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -n 6 $OUT/pytest_gpu.log
BENCH_STEPS=10 tools/gpu_try.sh r02j "" "default:--torch-gpu-rotations 0" "np32:lib=build/variants/libdlpd_np32.so --torch-gpu-rotations 0" \
  "nx128:lib=build/variants/libdlpd_nx128.so --torch-gpu-rotations 0" "epipk:lib=build/variants/libdlpd_epipk.so --torch-gpu-rotations 0" \
  "epidual:lib=build/variants/libdlpd_epidual.so --torch-gpu-rotations 0"
timeout 600 python bench.py --config 5 --steps 4 --warmup 3 --rot-batch 32 --max-rot 32 > $OUT/bench_cfg5.log 2>&1; echo "rc=$?" >> $OUT/bench_cfg5.log
timeout 600 python bench.py --config 4 --steps 8 --warmup 3 > $OUT/bench_cfg4.log 2>&1; echo "rc=$?" >> $OUT/bench_cfg4.log
python - <<PY
import json
for name in ("bench_cfg5", "bench_cfg4"):
    try:
        l=[x for x in open("$OUT/%s.log" % name) if x.startswith("{")][-1]; d=json.loads(l)
        print(name, "rot/s %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]), {k: round(v["ms_per_launch"],3) for k,v in d["kernels"].items()}, d["roofline_step"]["frac"])
    except Exception as e:
        print("failed", name, e); print(open("$OUT/%s.log" % name).read()[-1500:])
PY
