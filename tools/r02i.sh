python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for v in s4 s4keep s4k3 s8; do echo "== $v"; MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 1000000 --reps 5 | cut -c1-90; done
MGQP_B200_LIB=$PWD/variants/s4.so ncu --set full --clock-control none --import-source on -k regex:gi_lane -c 1 -o gpurun_out/prof_r02i_s4 -f python tools/prof_one.py --workload C2 --batch 1000000 --reps 1 > gpurun_out/r02i_ncu.log 2>&1
