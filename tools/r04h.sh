# full capture of the shipped lane kernel (unrolled set-up, grouped divisions) on C2 1M; timing first, without ncu
python tools/prof_one.py --workload C2 --batch 1000000 --reps 10 | cut -c1-100
ncu --set full --clock-control none --import-source on -k regex:gi_lane -c 1 -o gpurun_out/prof_r04h -f python tools/prof_one.py --workload C2 --batch 1000000 --reps 1 > gpurun_out/r04h_ncu.log 2>&1
tail -2 gpurun_out/r04h_ncu.log
