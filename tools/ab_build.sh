#!/bin/bash
# Build kernel variants for side-by-side timing on one GPU box.
# usage: tools/ab_build.sh name1="-DFOO=1 -DBAR" name2="..."   -> variants/<name>.so (git-ignored, travels with gpurun)
# then on the box:  for v in variants/*.so; do MGQP_B200_LIB=$PWD/$v python tools/prof_one.py ...; done
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
CS=motion-generation-using-quadratic-programs_b200/csrc
for spec in "$@"; do
  name=${spec%%=*}; flags=${spec#*=}
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -ffp-contract=off -Xcompiler -fPIC -shared \
      $flags -I include -o variants/$name.so $CS/mgqp_b200.cu $CS/quadprog_dropin.cc $CS/host_pack.cc -Xptxas -v 2>&1 \
      | grep -A2 "Compiling.*gi_lane_kernelILi7ELb0" | tail -1 | sed "s/^/$name: /" ) &
done
wait
ls -la variants
