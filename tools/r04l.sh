# A/B: L2 prefetch of the NEXT tile's CI / ci0 after phase B (pf1: in addition to this tile's prefetch in phase A; pf2: instead)
for rep in 1 2 3; do
for v in pf0 pf1 pf2; do
  MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 4000000 --reps 10 2>&1 | tail -1 | sed "s/^/$v: /"
done; done
