# A/B: add_constraint sweep one rotation at a time (ln_givens) vs split passes (hypotenuse chain, then the divisions of
# 1 / 2 / 3 / 6 rotations side by side)
for rep in 1 2; do
for v in sweep0 sweepg1 sweepg2 sweep2 sweepg6; do
  MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 4000000 --reps 10 2>&1 | tail -1 | sed "s/^/$v: /"
done; done
MGQP_B200_LIB=$PWD/variants/sweep2.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
