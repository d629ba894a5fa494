#!/bin/bash
# usage: tools/phase.sh <tag>   -- phase breakdown of gpurun_out/prof_<tag>.ncu-rep for the lane kernel (cubin must match the build)
T=$1; K=${2:-gi_lane_kernelILi7ELb0}
cd "$(dirname "$0")/.."
ncu -i gpurun_out/prof_$T.ncu-rep --page source --csv > gpurun_out/src_$T.csv 2>/dev/null
ncu -i gpurun_out/prof_$T.ncu-rep --page raw --csv > gpurun_out/raw_$T.csv 2>/dev/null
L=motion-generation-using-quadratic-programs_b200/csrc/gi_lane.cuh
b=($(grep -n "set-up$\|inequality columns: classify\|J moves into\|while (!done)\|step 2a\|step 2b\|add_constraint (App\|delete_constraint(del_l)\|if (degenerate) {" $L | cut -d: -f1))
g0=$(grep -n "MGQP_DEV double ln_distance" $L | cut -d: -f1)
python tools/ncu_by_phase.py gpurun_out/src_$T.csv /tmp/w/mgqp.cubin $K gi_lane.cuh givens:$g0-$((b[0]-30)) setup:$((b[0]-29))-$((b[1]-1)) classify:${b[1]}-$((b[2]-1)) jload:${b[2]}-$((b[3]-1)) scan:${b[3]}-$((b[4]-1)) direction:${b[4]}-$((b[5]-1)) step:${b[5]}-$((b[6]-1)) add:${b[6]}-$((b[7]-1)) delete:${b[7]}-$((b[8]-1)) tail:${b[8]}-$((b[8]+40))
