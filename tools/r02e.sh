python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for v in k3 k3sync k4 k2 k6; do echo "== $v"; MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 1000000 --reps 5 | cut -c1-90; done
for w in 4 6 7; do echo "== k3 W=$w"; MGQP_LANE_WARPS=$w MGQP_B200_LIB=$PWD/variants/k3.so python tools/prof_one.py --workload C2 --batch 1000000 --reps 5 | cut -c1-90; done
MGQP_B200_LIB=$PWD/variants/k3.so ncu --set full --clock-control none --import-source on -k regex:gi_lane -c 1 -o gpurun_out/prof_r02e_k3 -f python tools/prof_one.py --workload C2 --batch 1000000 --reps 1 > gpurun_out/r02e_ncu.log 2>&1
MGQP_B200_LIB=$PWD/variants/k3sync.so ncu --set full --clock-control none --import-source on -k regex:gi_lane -c 1 -o gpurun_out/prof_r02e_k3sync -f python tools/prof_one.py --workload C2 --batch 1000000 --reps 1 > gpurun_out/r02e_ncu2.log 2>&1
