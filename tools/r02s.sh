ncu --section SourceCounters --section WarpStateStats --section MemoryWorkloadAnalysis --section SpeedOfLight --clock-control none --import-source on -k regex:gi_large -c 1 -o gpurun_out/prof_r02s_large -f python tools/prof_c4.py 1776 > gpurun_out/r02s_ncu.log 2>&1
tail -2 gpurun_out/r02s_ncu.log
