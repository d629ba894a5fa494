# 3 CTAs of 256 threads per SM (row block 4) vs 2 x 384 (row block 8 / 4), exact and fast mode; parity of the 3-CTA build
run() { MGQP_B200_LIB=$PWD/variants/$1.so MGQP_LARGE_T_EXACT=$2 MGQP_LARGE_CTAS_PER_SM=$3 python - <<P
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
    print('$1 T=$2 ctas/sm=$3 flags',fl,'C4 2368: %.1f ms -> %.4g QP/s'%(dt*1e3, 2368/dt))
P
}
run rb4 256 3; run rb2 192 4; run rb3 224 3; run rb2 192 3
MGQP_B200_LIB=$PWD/variants/rb2.so MGQP_LARGE_T_EXACT=192 MGQP_LARGE_CTAS_PER_SM=4 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cta or mpc or fast_math" 2>&1 | tail -3
