# A/B: set-up divisions one per call vs grouped (2-3 quotients over one divisor per call)
for rep in 1 2 3; do
for v in div1 divg; do
  MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 4000000 --reps 10 2>&1 | tail -1 | sed "s/^/$v: /"
done; done
MGQP_B200_LIB=$PWD/variants/divg.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane or headline or C2 or C5 or bad or poison" 2>&1 | tail -3
