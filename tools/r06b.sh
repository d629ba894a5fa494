# four CTAs per SM (49 KB each): 128 / 160 / 192 threads
run() { MGQP_B200_LIB=$PWD/variants/$1.so MGQP_LARGE_T_EXACT=$2 python - <<P
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
    print('$1 T=$2 flags',fl,'C4 2368: %.1f ms -> %.4g QP/s'%(dt*1e3, 2368/dt))
P
}
run d128 96; run d128 64
