#!/bin/bash
# Build variants of libdlpd.so locally (they travel to the GPU box with the snapshot): build/variants/libdlpd_<name>.so
# usage: tools/build_variants.sh "name1:-DX=1 -DY=2" "name2:..."     (up to 4 builds in parallel)
set -u
mkdir -p build/variants
n=0
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  ( nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared $flags -Xptxas -v \
      -o build/variants/libdlpd_$name.so deeplocalproteindocking_b200/csrc/dlpd_api.cu > build/variants/build_$name.log 2>&1 \
      && echo "built $name" || echo "FAILED $name" ) &
  n=$((n+1)); if [ $((n % 4)) -eq 0 ]; then wait; fi
done
wait
