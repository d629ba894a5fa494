./variants/tmem_probe
run() { echo "W=$1 C=$2 carve=$3"; env MGQP_LANE_WARPS=$1 MGQP_LANE_CTAS_PER_SM=$2 MGQP_LANE_CARVEOUT=$3 python tools/prof_one.py --workload C2 --batch 1000000 --reps 5 | cut -c1-80; }
for cv in 60 70 100; do run 4 1 $cv; done
for cv in 80 100; do run 5 1 $cv; done
