# A/B: CTA-per-QP kernel, hypotenuse chain of add_constraint with the builtin division / sqrt vs their split inline paths
for v in lc0 lc1 lc0 lc1; do
MGQP_B200_LIB=$PWD/variants/$v.so python - <<P
import sys, time, os; sys.path.insert(0,'.')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
dev=torch.device('cuda:0')
t=workloads.make_c4_device(2368, dev)
for fl in (0,1):
    for rep in range(3):
        torch.cuda.synchronize(); t0=time.time()
        fa=mgqp_b200.solve_quadprog_batched(*t, flags=fl); torch.cuda.synchronize()
        dt=time.time()-t0
    print('$v flags',fl,'C4 2368: %.1f ms -> %.4g QP/s'%(dt*1e3, 2368/dt), 'iters mean', float(fa.status.float().mean()))
P
done
MGQP_B200_LIB=$PWD/variants/lc1.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cta or mpc" 2>&1 | tail -3
