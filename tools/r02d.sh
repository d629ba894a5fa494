python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { echo "W=$1 C=$2"; env MGQP_LANE_WARPS=$1 MGQP_LANE_CTAS_PER_SM=$2 python tools/prof_one.py --workload ${3:-C2} --batch 1000000 --reps 5 | cut -c1-100; }
for w in 4 5 6 7 8; do run $w 1; done
run 8 1 C5
ncu --set full --clock-control none --import-source on -k regex:gi_lane -c 1 -o gpurun_out/prof_r02d_lane -f python tools/prof_one.py --workload C2 --batch 1000000 --reps 1 > gpurun_out/r02d_ncu.log 2>&1
