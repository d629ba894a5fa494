python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for v in k3 k3sync k2sync k4sync; do echo "== $v"; MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 1000000 --reps 5 | cut -c1-90; done
echo "== k3sync C5"; MGQP_B200_LIB=$PWD/variants/k3sync.so python tools/prof_one.py --workload C5 --batch 1000000 --reps 5 | cut -c1-90
MGQP_B200_LIB=$PWD/variants/k3sync.so ncu --set full --clock-control none --import-source on -k regex:gi_lane -c 1 -o gpurun_out/prof_r02f_k3sync -f python tools/prof_one.py --workload C2 --batch 1000000 --reps 1 > gpurun_out/r02f_ncu.log 2>&1
