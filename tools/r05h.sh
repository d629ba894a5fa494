# A/B: group kernel's hypotenuse chain on the split division / square root (gd) vs HEAD
for rep in 1 2; do
for v in head gd; do
  for w in S0:400000 C3:100000 S2:400000; do
  MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload ${w%%:*} --batch ${w##*:} --reps 10 2>&1 | tail -1 | cut -c1-70 | sed "s/^/$v: /"
  done
done; done
MGQP_B200_LIB=$PWD/variants/gd.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
