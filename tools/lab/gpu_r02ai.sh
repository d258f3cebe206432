#!/bin/bash
# round 2, visit ai: ncu launch list + full captures of the search kernels on the FINAL build (config 2), and a full capture of the
# 128^3 kernels in their new shapes (config 5)
set -u
OUT=gpurun_out/r02ai; mkdir -p $OUT
ARGS="--steps 1 --warmup 1 --rot-batch 32 --max-rot 32 --cpu-rotations 0 --torch-gpu-rotations 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dlpd_kernel -c 80 --csv \
  --log-file $OUT/launches.csv python bench.py $ARGS > $OUT/launches_run.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k "regex:Epilogue|Plane|FwdZC|HiddenL1" -s 12 -c 6 -o $OUT/prof -f python bench.py $ARGS > $OUT/ncu_run.log 2>&1
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
python tools/ncu_summary.py $OUT/raw.csv > $OUT/ncu_full_summary.txt 2>&1
cut -c1-200 $OUT/ncu_full_summary.txt | head -12
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k "regex:Epilogue|Plane|FwdZC|HiddenL1" -s 12 -c 6 -o $OUT/prof_cfg5 -f python bench.py --config 5 $ARGS > $OUT/ncu_run_cfg5.log 2>&1
ncu -i $OUT/prof_cfg5.ncu-rep --page raw --csv > $OUT/raw_cfg5.csv 2>/dev/null
python tools/ncu_summary.py $OUT/raw_cfg5.csv > $OUT/ncu_full_summary_cfg5.txt 2>&1
cut -c1-200 $OUT/ncu_full_summary_cfg5.txt | head -12
rm -f $OUT/prof_cfg5.ncu-rep
ls -la $OUT
