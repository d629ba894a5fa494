python - <<'P'
import sys, time, os; sys.path.insert(0,'.')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
dev=torch.device('cuda:0')
t=workloads.make_c4_device(2368, dev)
for T,c in (('256','2'),('128','4'),('128','3'),('256','3'),('64','6'),('512','1')):
    os.environ['MGQP_LARGE_T']=T; os.environ['MGQP_LARGE_CTAS_PER_SM']=c
    for fl in (0,1):
        for rep in range(3):
            torch.cuda.synchronize(); t0=time.time()
            out=mgqp_b200.solve_quadprog_batched(*t, flags=fl); torch.cuda.synchronize()
            dt=time.time()-t0
        print('T',T,'ctas/sm',c,'flags',fl,'C4 2368: %.1f ms -> %.4g QP/s'%(dt*1e3, 2368/dt))
P
