#!/bin/bash
# round 2, visit z: parity sweeps on the inputs of configs 3 / 4 / 5 (tests/sweeps/parity_sweep.py), then the round-end sequence
set -u
OUT=gpurun_out/r02z; mkdir -p $OUT
timeout 600 python tests/sweeps/parity_sweep.py --config 4 --rotations 128 --pairs 4 > $OUT/parity_sweep_cfg4.log 2>&1; echo "rc=$?" >> $OUT/parity_sweep_cfg4.log; tail -2 $OUT/parity_sweep_cfg4.log
timeout 600 python tests/sweeps/parity_sweep.py --config 3 --rotations 128 > $OUT/parity_sweep_cfg3.log 2>&1; echo "rc=$?" >> $OUT/parity_sweep_cfg3.log; tail -2 $OUT/parity_sweep_cfg3.log
timeout 900 python tests/sweeps/parity_sweep.py --config 5 --rotations 32 > $OUT/parity_sweep_cfg5.log 2>&1; echo "rc=$?" >> $OUT/parity_sweep_cfg5.log; tail -2 $OUT/parity_sweep_cfg5.log
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log; tail -2 $OUT/smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > $OUT/bench.log 2>&1; echo "bench rc=$?" >> $OUT/bench.log
python - <<PY
import json
for name in ("bench",):
    try:
        l=[x for x in open("$OUT/%s.log" % name) if x.startswith("{")][-1]; d=json.loads(l)
        print(name, "value %.2f e2e %.2f steps %d ms/step %.1f" % (d["value"], d["e2e"]["value"], d["steps"], d["ms_per_step"]), d.get("parity"), d.get("clocks"))
    except Exception as e:
        print("failed", name, e); print(open("$OUT/%s.log" % name).read()[-1500:])
PY
