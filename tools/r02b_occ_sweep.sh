# occupancy / CTA-shape / carve-out sweep of the lane kernel on C2 (1M QPs)
run() { echo "W=$1 C=$2 carve=$3"; MGQP_LANE_WARPS=$1 MGQP_LANE_CTAS_PER_SM=$2 ${3:+MGQP_LANE_CARVEOUT=$3} python tools/prof_one.py --workload C2 --batch 1000000 --reps 5 | cut -c1-80; }
nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 200 > gpurun_out/r02b_clocks.log &
SMI=$!
for c in 1 2 3 4 5; do run 1 $c; done
for c in 1 2; do run 2 $c; done
run 3 1; run 4 1; run 5 1; run 6 1
for cv in 50 66 75 86 100; do run 4 1 $cv; done
for cv in 75 86 100; do run 5 1 $cv; done
for cv in 50 100; do run 3 1 $cv; done
kill $SMI
sort gpurun_out/r02b_clocks.log | uniq -c | sort -rn | head -5
for w in 3 5 6; do
MGQP_LANE_WARPS=$w ncu --set full --clock-control none --import-source on -k regex:gi_lane -c 1 -o gpurun_out/prof_r02b_w$w -f python tools/prof_one.py --workload C2 --batch 1000000 --reps 1 > gpurun_out/r02b_ncu_w$w.log 2>&1
done
