# A/B: rolled vs unrolled lane set-up, same box, interleaved, 4M C2 QPs; then lane parity tests on the unrolled build
for rep in 1 2 3; do
for v in rolled unrolled; do
  MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 4000000 --reps 10 2>&1 | tail -1 | sed "s/^/$v: /"
done; done
for v in rolled unrolled; do
  MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C5 --batch 4000000 --reps 10 2>&1 | tail -1 | sed "s/^/$v: /"
done
MGQP_B200_LIB=$PWD/variants/unrolled.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
