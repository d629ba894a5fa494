python - <<'P'
import sys, time, os; sys.path.insert(0,'.')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
dev=torch.device('cuda:0')
t=workloads.make_c4_device(1184, dev)
for T in ('256','512','1024'):
  os.environ['MGQP_LARGE_T']=T
  for c in ('32','48','64','84','110','148','296'):
    os.environ['MGQP_LARGE_MAX_CTAS']=c
    for rep in range(2):
        torch.cuda.synchronize(); t0=time.time()
        out=mgqp_b200.solve_quadprog_batched(*t, flags=1); torch.cuda.synchronize()
        dt=time.time()-t0
    print('T',T,'max ctas',c,'C4 1184 fast: %.1f ms -> %.4g QP/s'%(dt*1e3, 1184/dt))
P
