# A/B: ln_givens with the builtin divisions vs the split inline path (branch-free, c and s share 1/h)
for rep in 1 2 3; do
for v in giv0 giv1; do
  MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 4000000 --reps 10 2>&1 | tail -1 | sed "s/^/$v: /"
done; done
MGQP_B200_LIB=$PWD/variants/giv1.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
