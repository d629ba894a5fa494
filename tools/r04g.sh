# A/B: set-up divisions through grouped calls (head) vs the split inline path with one refined reciprocal per pivot
for rep in 1 2 3; do
for v in head sdiv; do
  MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 4000000 --reps 10 2>&1 | tail -1 | sed "s/^/$v: /"
done; done
MGQP_B200_LIB=$PWD/variants/sdiv.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
