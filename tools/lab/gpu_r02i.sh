#!/bin/bash
# round 2, visit i: ncu launch list + full captures of the search kernels (tcgen05 epilogue)
set -u
OUT=gpurun_out/r02i; mkdir -p $OUT
ARGS="--steps 1 --warmup 1 --rot-batch 32 --max-rot 32 --cpu-rotations 0 --torch-gpu-rotations 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dlpd_kernel -c 80 --csv \
  --log-file $OUT/launches.csv python bench.py $ARGS > $OUT/launches_run.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k "regex:Epilogue|Plane|FwdZC|HiddenL1" -s 12 -c 6 -o $OUT/prof -f python bench.py $ARGS > $OUT/ncu_run.log 2>&1
ls -la $OUT
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/raw.csv > $OUT/ncu_full_summary.txt 2>&1
cat $OUT/ncu_full_summary.txt | cut -c1-220
