for v in base parhyp nosd; do
MGQP_B200_LIB=$PWD/variants/$v.so python - <<P
import sys, time, os; sys.path.insert(0,'.')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
dev=torch.device('cuda:0')
t=workloads.make_c4_device(2368, dev)
res={}
for fl in (0,1):
    for rep in range(3):
        torch.cuda.synchronize(); t0=time.time()
        out=mgqp_b200.solve_quadprog_batched(*t, flags=fl, want_active_set=True); torch.cuda.synchronize()
        dt=time.time()-t0
    res[fl]=out
    print('$v flags',fl,'C4 2368: %.1f ms -> %.4g QP/s'%(dt*1e3, 2368/dt))
ex,fa=res[0],res[1]
print('$v A equal', float((ex.A==fa.A).all(dim=1).double().mean()), 'x rel', float(((ex.x-fa.x).abs().amax(dim=1)/ex.x.abs().amax(dim=1).clamp(min=1)).max()))
z=lambda *sh: torch.zeros(sh, dtype=torch.float64, device=dev)
n=350; B=1184
bare=[t[0][:B], t[1][:B], z(B,n,0), z(B,0), z(B,n,0), z(B,0)]
for fl in (16,17):
    for rep in range(3):
        torch.cuda.synchronize(); t0=time.time(); o=mgqp_b200.solve_quadprog_batched(*bare, flags=fl); torch.cuda.synchronize(); dt=time.time()-t0
    print('$v dense part flags',fl,'%.2f ms'%(dt*1e3))
P
done
