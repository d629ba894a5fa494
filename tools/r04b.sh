# A/B: tiles per warp between two group barriers (1/2/3), group size 8 / 2 with the unrolled set-up
for rep in 1 2; do
for v in sr1 sr2 sr3 gw8 gw2; do
  MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 4000000 --reps 10 2>&1 | tail -1 | sed "s/^/$v: /"
done; done
