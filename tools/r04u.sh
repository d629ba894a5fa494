# C4 (CTA-per-QP kernel), exact mode, current build: source-level capture for the phase split + iteration statistics
python - <<'P'
import sys; sys.path.insert(0,'.')
import torch, mgqp_b200
from mgqp_b200 import workloads
t = workloads.make_c4_device(1184, torch.device("cuda:0"))
out = mgqp_b200.solve_quadprog_batched(*t, flags=0, want_active_set=True)
torch.cuda.synchronize()
it = out.iters.float(); iq = out.iq.float()
print('C4 1184: iters mean %.1f max %d; final |A| mean %.1f max %d; status %s' % (it.mean(), int(it.max()), iq.mean(), int(iq.max()), out.status.bincount().tolist()))
P
ncu --section SourceCounters --section WarpStateStats --section MemoryWorkloadAnalysis --section SpeedOfLight --clock-control none --import-source on -k regex:gi_large -c 1 -o gpurun_out/prof_r04u_large -f python tools/prof_c4.py 1776 0 > gpurun_out/r04u_ncu.log 2>&1
tail -2 gpurun_out/r04u_ncu.log
