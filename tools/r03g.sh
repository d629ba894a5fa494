nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv,noheader
python - <<'P'
import sys, time, os; sys.path.insert(0,'.')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
dev=torch.device('cuda:0')
t=workloads.make_c4_device(2368, dev)
for was in (False, True):
  for fl in (0,1):
    for rep in range(3):
        torch.cuda.synchronize(); t0=time.time()
        out=mgqp_b200.solve_quadprog_batched(*t, flags=fl, want_active_set=was); torch.cuda.synchronize()
        dt=time.time()-t0
    print('active_set',was,'flags',fl,'C4 2368: %.1f ms -> %.4g QP/s'%(dt*1e3, 2368/dt))
P
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv,noheader
