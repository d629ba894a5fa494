python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for v in s8 s4 s2 s8k3; do echo "== $v"; MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 1000000 --reps 5 | cut -c1-90; done
MGQP_B200_LIB=$PWD/variants/s8.so ncu --set full --clock-control none --import-source on -k regex:gi_lane -c 1 -o gpurun_out/prof_r02g_s8 -f python tools/prof_one.py --workload C2 --batch 1000000 --reps 1 > gpurun_out/r02g_ncu.log 2>&1
