python - <<'P'
import sys, time, os; sys.path.insert(0,'.')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
dev=torch.device('cuda:0')
t=workloads.make_c4_device(2368, dev)
z=lambda *sh: torch.zeros(sh, dtype=torch.float64, device=dev)
n=350; B=1184
bare=[t[0][:B], t[1][:B], z(B,n,0), z(B,0), z(B,n,0), z(B,0)]
for gate in ('0','16','32','64','96'):
    os.environ['MGQP_LARGE_SETUP_GATE']=gate
    for rep in range(2):
        torch.cuda.synchronize(); t0=time.time()
        fa=mgqp_b200.solve_quadprog_batched(*t, flags=1); torch.cuda.synchronize()
        dt=time.time()-t0
    o=mgqp_b200.solve_quadprog_batched(*bare, flags=17); torch.cuda.synchronize()
    t0=time.time(); o=mgqp_b200.solve_quadprog_batched(*bare, flags=17); torch.cuda.synchronize(); dd=time.time()-t0
    print('gate',gate,'C4 fast: %.1f ms -> %.4g QP/s; dense part %.1f ms'%(dt*1e3, 2368/dt, dd*1e3))
for cps in ('1',):
    os.environ['MGQP_LARGE_SETUP_GATE']='0'; os.environ['MGQP_LARGE_CTAS_PER_SM']=cps
    for rep in range(2):
        torch.cuda.synchronize(); t0=time.time()
        fa=mgqp_b200.solve_quadprog_batched(*t, flags=1); torch.cuda.synchronize()
        dt=time.time()-t0
    print('ctas/sm',cps,'C4 fast: %.1f ms -> %.4g QP/s'%(dt*1e3, 2368/dt))
P
