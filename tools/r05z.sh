# x_old / u_old out of shared memory: five CTAs of 128 threads really fit an SM now (43 KB each)
python - <<'P'
import sys, time, os; sys.path.insert(0,'.')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
dev=torch.device('cuda:0')
t=workloads.make_c4_device(2368, dev)
for cap in ('4','5'):
    os.environ['MGQP_LARGE_CTAS_PER_SM']=cap
    for fl in (0,1):
        for rep in range(3):
            torch.cuda.synchronize(); t0=time.time()
            fa=mgqp_b200.solve_quadprog_batched(*t, flags=fl); torch.cuda.synchronize()
            dt=time.time()-t0
        print('ctas/sm',cap,'flags',fl,'C4 2368: %.1f ms -> %.4g QP/s'%(dt*1e3, 2368/dt))
P
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cta or mpc or fast_math or degenerate" 2>&1 | tail -3
