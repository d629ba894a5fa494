# A/B: phase A of the next tile issued before the round's barrier (sa1) vs at the start of the tile (sa0)
for rep in 1 2 3; do
for v in sa0 sa1; do
  MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 4000000 --reps 10 2>&1 | tail -1 | sed "s/^/$v: /"
done; done
for v in sa0 sa1; do MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C5 --batch 4000000 --reps 10 2>&1 | tail -1 | sed "s/^/$v: /"; done
MGQP_B200_LIB=$PWD/variants/sa1.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
