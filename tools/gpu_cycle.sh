#!/bin/bash
# One GPU round trip for the lane kernel: parity tests, quick timing, ncu capture.  usage: tools/gpu_cycle.sh <tag> [workload] [batch]
TAG=$1; WL=${2:-C2}; B=${3:-1000000}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${TAG}_tests.log
python tools/prof_one.py --workload $WL --batch $B --reps 5 > gpurun_out/${TAG}_time.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gi_ -c 1 -o gpurun_out/prof_${TAG} -f python tools/prof_one.py --workload $WL --batch $B --reps 1 > gpurun_out/${TAG}_ncu.log 2>&1
cat gpurun_out/${TAG}_tests.log gpurun_out/${TAG}_time.log
