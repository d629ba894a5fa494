python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python - <<'P'
import sys, time, os; sys.path.insert(0,'.')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
dev=torch.device('cuda:0')
t=workloads.make_c4_device(2368, dev)
for rep in range(3):
    torch.cuda.synchronize(); t0=time.time()
    out=mgqp_b200.solve_quadprog_batched(*t); torch.cuda.synchronize()
    dt=time.time()-t0
    print('C4 2368 QPs: %.1f ms -> %.4g QP/s'%(dt*1e3, 2368/dt))
for T in ('128','512'):
    os.environ['MGQP_LARGE_T']=T
    for rep in range(2):
        torch.cuda.synchronize(); t0=time.time()
        out=mgqp_b200.solve_quadprog_batched(*t); torch.cuda.synchronize()
        dt=time.time()-t0
    print('T',T,'C4 2368 QPs: %.1f ms -> %.4g QP/s'%(dt*1e3, 2368/dt))
P
