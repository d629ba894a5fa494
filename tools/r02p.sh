python - <<'P'
import sys, time, os, subprocess; sys.path.insert(0,'.')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
dev=torch.device('cuda:0')
t=workloads.make_c4_device(2368, dev)
ref=None
for gate in ('0','8','16','32','64','128'):
    os.environ['MGQP_LARGE_SETUP_GATE']=gate
    for rep in range(2):
        torch.cuda.synchronize(); t0=time.time()
        out=mgqp_b200.solve_quadprog_batched(*t); torch.cuda.synchronize()
        dt=time.time()-t0
    if ref is None: ref=out.x.clone()
    print('gate',gate,'C4 2368 QPs: %.1f ms -> %.4g QP/s'%(dt*1e3, 2368/dt),'same',bool(torch.equal(out.x,ref)))
P
