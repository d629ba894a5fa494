# A/B: ln_givens per rotation vs ln_givens_next (rotation j's coefficients + rotation j-1's hypotenuse in one call)
for rep in 1 2 3; do
for v in chain0 chain1; do
  MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 4000000 --reps 10 2>&1 | tail -1 | sed "s/^/$v: /"
done; done
MGQP_B200_LIB=$PWD/variants/chain1.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
