# A/B: phase B's column byte computed (ct0) vs looked up in a 4<<N-byte shared-memory table (ct1)
for rep in 1 2 3; do
for v in ct0 ct1; do
  MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 4000000 --reps 10 2>&1 | tail -1 | sed "s/^/$v: /"
done; done
MGQP_B200_LIB=$PWD/variants/ct1.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
