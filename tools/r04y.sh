# A/B: phase B with the shared-window base hoisted (pb0), plus the GEN flag settled after the loop (pb1), vs HEAD
for rep in 1 2 3; do
for v in head pb0 pb1; do
  MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 4000000 --reps 10 2>&1 | tail -1 | sed "s/^/$v: /"
done; done
MGQP_B200_LIB=$PWD/variants/pb1.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
