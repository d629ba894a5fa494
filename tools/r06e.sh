# full capture of the group kernel on S0 with the inequality block handed over (and kept) column-compressed
T=r06e
timeout 150 ncu --set full --clock-control none --import-source on -k regex:gi_small -c 1 -o gpurun_out/prof_${T}_s0c -f \
  python tools/prof_one.py --workload S0 --batch 200000 --reps 1 --compressed > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
