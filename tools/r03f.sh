python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<'P'
import sys, time, os; sys.path.insert(0,'.')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
dev=torch.device('cuda:0')
t=workloads.make_c4_device(2368, dev)
for fl in (0,1):
    for rep in range(2):
        torch.cuda.synchronize(); t0=time.time()
        out=mgqp_b200.solve_quadprog_batched(*t, flags=fl, want_active_set=True); torch.cuda.synchronize()
        dt=time.time()-t0
    print('flags',fl,'C4 2368: %.1f ms -> %.4g QP/s'%(dt*1e3, 2368/dt))
    if fl==0: ex=out
    else: print('A equal', float((ex.A==out.A).all(dim=1).double().mean()), 'x rel', float(((ex.x-out.x).abs().amax(dim=1)/ex.x.abs().amax(dim=1).clamp(min=1)).max()))
P
