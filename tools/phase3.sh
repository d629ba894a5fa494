#!/bin/bash
# usage: tools/phase3.sh <tag> <cubin>  -- phase breakdown of gpurun_out/prof_<tag>.ncu-rep (lane kernel v5)
T=$1; CUBIN=$2; K=${3:-gi_lane_kernelILi7ELb0}
cd "$(dirname "$0")/.."
ncu -i gpurun_out/prof_$T.ncu-rep --page source --csv > gpurun_out/src_$T.csv 2>/dev/null
ncu -i gpurun_out/prof_$T.ncu-rep --page raw --csv > gpurun_out/raw_$T.csv 2>/dev/null
python tools/ncu_by_phase.py gpurun_out/src_$T.csv $CUBIN $K gi_lane.cuh $(bash tools/lane_ranges.sh)
