ncu --set full --clock-control none --import-source on -k regex:gi_large -c 1 -o gpurun_out/prof_r02q_large -f python tools/prof_one.py --workload C4 --batch 592 --reps 1 > gpurun_out/r02q_ncu.log 2>&1
tail -2 gpurun_out/r02q_ncu.log
