ncu --section SourceCounters --section WarpStateStats --clock-control none --import-source on -k regex:gi_small -c 1 -o gpurun_out/prof_r05i_s0 -f python tools/prof_one.py --workload S0 --batch 200000 --reps 1 > gpurun_out/r05i_ncu.log 2>&1
tail -2 gpurun_out/r05i_ncu.log
