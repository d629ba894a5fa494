# A/B: ci0 into the slot by cp.async (c0a) vs loaded with the column's rows and stored (c0b); both with phase A's hoisted base
for rep in 1 2 3; do
for v in c0a c0b; do
  MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 4000000 --reps 10 2>&1 | tail -1 | sed "s/^/$v: /"
done; done
MGQP_B200_LIB=$PWD/variants/c0b.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "benchmark_shapes or random_shapes or tail or non_finite or broadcast or column_compressed" 2>&1 | tail -3
