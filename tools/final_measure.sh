#!/bin/bash
# One GPU call that produces everything profiles/ cites: bench lines for every workload, the reference arm,
# the ncu launch list (with DRAM bytes) of the default bench command, and a --set full capture of the headline kernel.
# usage: tools/final_measure.sh <tag>      (outputs under gpurun_out/<tag>_*)
T=$1
python bench.py > gpurun_out/${T}_bench_C2.json 2> gpurun_out/${T}_bench_C2.err
for w in C5 C3 S0 S2; do python bench.py --workload $w --steps 5 > gpurun_out/${T}_bench_$w.json 2>/dev/null; done
python bench.py --workload C4 --steps 2 --e2e-steps 1 > gpurun_out/${T}_bench_C4.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/${T}_launches_bench_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gi_lane -c 1 -o gpurun_out/prof_${T}_lane -f \
  python tools/prof_one.py --workload C2 --batch 1000000 --reps 1 > gpurun_out/${T}_ncu_full.log 2>&1
for w in C2 C5 C3 S0 S2 C4 ref; do echo "== $w"; cat gpurun_out/${T}_bench_$w.json | cut -c1-330; done
