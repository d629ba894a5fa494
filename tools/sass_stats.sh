#!/bin/bash
# usage: tools/sass_stats.sh [kernel-substring]   -- compile the product TU to a cubin and print ptxas + SASS size info
K=${1:-gi_lane_kernelILi7ELb0}
mkdir -p /tmp/w
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -I include -cubin -o /tmp/w/mgqp.cubin \
  motion-generation-using-quadratic-programs_b200/csrc/mgqp_b200.cu -Xptxas -v 2>&1 | grep -A2 "Compiling.*$K" | tail -2
cuobjdump -sass /tmp/w/mgqp.cubin | awk -v k="$K" '/Function : /{f=($3 ~ k)} f && /^\s+\/\*[0-9a-f]{4}\*\//{c++} END{print c, "SASS instructions"}'
nvdisasm --print-line-info -c /tmp/w/mgqp.cubin > /tmp/w/dis.txt 2>/dev/null
