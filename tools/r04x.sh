# A/B: scan over "effective bytes" (index of the column's term or of +inf) vs index mask + flag test
for rep in 1 2 3; do
for v in eff0 eff1; do
  MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 4000000 --reps 10 2>&1 | tail -1 | sed "s/^/$v: /"
done; done
for v in eff0 eff1; do MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C5 --batch 4000000 --reps 10 2>&1 | tail -1 | sed "s/^/$v: /"; done
MGQP_B200_LIB=$PWD/variants/eff1.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
