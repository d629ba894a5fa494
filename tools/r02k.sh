python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in s4 s2 s8; do echo "== $v"; MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 1000000 --reps 5 | cut -c1-90; done
for w in C5 S0 S2 C3; do python tools/prof_one.py --workload $w --batch 400000 --reps 3 | cut -c1-100; done
