ncu --section SourceCounters --section WarpStateStats --section MemoryWorkloadAnalysis --section SpeedOfLight --clock-control none --import-source on -k regex:gi_large -c 1 -o gpurun_out/prof_r02v_large_fast -f python tools/prof_c4.py 1776 1 > gpurun_out/r02v_ncu.log 2>&1
tail -2 gpurun_out/r02v_ncu.log
