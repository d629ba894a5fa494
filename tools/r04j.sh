# A/B: builtin sqrt vs split sqrt inside ln_givens (c0s0 / c0s1); ln_givens_next with everything branch-free (c1s1)
for rep in 1 2 3; do
for v in c0s0 c0s1 c1s1; do
  MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 4000000 --reps 10 2>&1 | tail -1 | sed "s/^/$v: /"
done; done
MGQP_B200_LIB=$PWD/variants/c1s1.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split or sparse or benchmark_shapes or random_shapes or non_finite or degenerate" 2>&1 | tail -3
MGQP_B200_LIB=$PWD/variants/c0s1.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "split or sparse or benchmark_shapes or random_shapes or non_finite or degenerate" 2>&1 | tail -3
