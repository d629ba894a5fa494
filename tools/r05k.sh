# A/B: unroll depth of the exact-mode set-up loops of the CTA-per-QP kernel (loads in flight per thread)
for v in us4 us8 us16 us32 us4 us8 us16 us32; do
MGQP_B200_LIB=$PWD/variants/$v.so python - <<P
import sys, time, os; sys.path.insert(0,'.')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
dev=torch.device('cuda:0')
t=workloads.make_c4_device(2368, dev)
for rep in range(3):
    torch.cuda.synchronize(); t0=time.time()
    fa=mgqp_b200.solve_quadprog_batched(*t, flags=0); torch.cuda.synchronize()
    dt=time.time()-t0
print('$v exact C4 2368: %.1f ms -> %.4g QP/s'%(dt*1e3, 2368/dt))
P
done
