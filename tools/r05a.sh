# A/B: R zeroing unrolled, running offsets into the packed R (update_r, column store) vs HEAD
for rep in 1 2 3; do
for v in head rh; do
  MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 4000000 --reps 10 2>&1 | tail -1 | sed "s/^/$v: /"
done; done
MGQP_B200_LIB=$PWD/variants/rh.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "benchmark_shapes or random_shapes or tail or non_finite or degenerate or failure" 2>&1 | tail -3
