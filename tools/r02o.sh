python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python - <<'P'
import sys, time; sys.path.insert(0,'.')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
dev=torch.device('cuda:0')
t=workloads.make_c4_device(2368, dev)
for rep in range(2):
    torch.cuda.synchronize(); t0=time.time()
    out=mgqp_b200.solve_quadprog_batched(*t, want_active_set=True); torch.cuda.synchronize()
    dt=time.time()-t0
    print('C4 2368 QPs: %.1f ms -> %.4g QP/s; mean steps %.1f'%(dt*1e3, 2368/dt, out.iters.double().mean().item()))
# broadcast CI (one instance for the batch)
t2=[t[0],t[1],t[2],t[3],t[4][:1].contiguous(),t[5]]
for rep in range(2):
    torch.cuda.synchronize(); t0=time.time()
    out2=mgqp_b200.solve_quadprog_batched(*t2); torch.cuda.synchronize()
    dt=time.time()-t0
    print('C4 shared CI: %.1f ms -> %.4g QP/s'%(dt*1e3, 2368/dt), 'same', bool(torch.equal(out2.x,out.x)))
P
