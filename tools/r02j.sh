for v in p1 p2 p3; do echo "== $v"; MGQP_B200_LIB=$PWD/variants/$v.so python tools/prof_one.py --workload C2 --batch 1000000 --reps 5 | cut -c1-90; done
