#!/bin/bash
# usage: tools/phase2.sh <tag>  -- phase breakdown of gpurun_out/prof_<tag>.ncu-rep for the lane kernel v5 (cubin /tmp/w/mgqp.cubin must match the build)
T=$1; K=${2:-gi_lane_kernelILi7ELb0}
cd "$(dirname "$0")/.."
ncu -i gpurun_out/prof_$T.ncu-rep --page source --csv > gpurun_out/src_$T.csv 2>/dev/null
ncu -i gpurun_out/prof_$T.ncu-rep --page raw --csv > gpurun_out/raw_$T.csv 2>/dev/null
L=motion-generation-using-quadratic-programs_b200/csrc/gi_lane.cuh
ln() { grep -n "$1" $L | head -1 | cut -d: -f1; }
g0=$(ln "MGQP_DEV double ln_distance"); sA=$(ln "^MGQP_DEV void ln_stage_G"); cb=$(ln "^MGQP_DEV unsigned ln_column_byte"); sB=$(ln "^MGQP_DEV void ln_stage_CI"); t0=$(ln "^MGQP_DEV void gi_lane_tile")
pA=$(ln "phase A: stage G, g0"); su=$(ln "set-up$"); pB=$(ln "phase B: inequality columns"); lp=$(ln "One pass of this loop"); sc=$(ln "step 1 / 2: scan"); d2=$(ln "step 2a"); s2=$(ln "step 2b"); ad=$(ln "add_constraint (App"); de=$(ln "delete_constraint(del_l)"); dg=$(ln "if (degenerate) {")
python tools/ncu_by_phase.py gpurun_out/src_$T.csv /tmp/w/mgqp.cubin $K gi_lane.cuh givens:$g0-$((sA-8)) stageG:$((sA-7))-$((cb-5)) stageCI:$((cb-4))-$((t0-3)) hdr:$((t0-2))-$((pA-1)) phaseA:$pA-$((su-1)) setup:$su-$((pB-1)) phaseB:$pB-$((lp-12)) init:$((lp-11))-$((sc-8)) scan:$((sc-7))-$((d2-9)) direction:$((d2-8))-$((s2-1)) step:$s2-$((ad-1)) add:$ad-$((de-1)) delete:$de-$((dg-1)) tail:$dg-$((dg+40))
