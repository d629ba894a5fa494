# A/B: CTA-per-QP kernel with 256 threads x 2 CTAs/SM (lb512) vs 352 threads (one row of the 350-variable shape per
# thread) x 2 CTAs/SM under __launch_bounds__(352, 2) = 80 registers (lb352); also 352 x 1 with 128 registers
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
run lb384 384; run lb416 416; run lb448 448; run lb384 384
MGQP_B200_LIB=$PWD/variants/lb384.so MGQP_LARGE_T_EXACT=384 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cta or mpc" 2>&1 | tail -3
