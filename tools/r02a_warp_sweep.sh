nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader; nproc
for w in 1 2 3 4 5 6; do echo "warps=$w"; MGQP_LANE_WARPS=$w python tools/prof_one.py --workload C2 --batch 1000000 --reps 5; done
