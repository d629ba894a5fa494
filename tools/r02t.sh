python - <<'P'
import sys, time, os; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
dev=torch.device('cuda:0')
t=workloads.make_c4_device(2368, dev)
ex=mgqp_b200.solve_quadprog_batched(*t, want_active_set=True); torch.cuda.synchronize()
for rep in range(2):
    torch.cuda.synchronize(); t0=time.time()
    fa=mgqp_b200.solve_quadprog_batched(*t, want_active_set=True, flags=1); torch.cuda.synchronize()
    dt=time.time()-t0
    print('C4 fast: %.1f ms -> %.4g QP/s'%(dt*1e3, 2368/dt))
print('status equal', bool(torch.equal(ex.status, fa.status)), 'iq equal', bool(torch.equal(ex.iq, fa.iq)), 'A equal frac', float((ex.A==fa.A).all(dim=1).double().mean()))
rel=((ex.x-fa.x).abs().amax(dim=1)/ex.x.abs().amax(dim=1).clamp(min=1.0))
print('x rel max', float(rel.max()), 'f rel max', float(((ex.f-fa.f).abs()/ex.f.abs().clamp(min=1.0)).max()), 'iters equal frac', float((ex.iters==fa.iters).double().mean()))
# dense part only
z=lambda *sh: torch.zeros(sh, dtype=torch.float64, device=dev)
n=350; B=1184
bare=[t[0][:B], t[1][:B], z(B,n,0), z(B,0), z(B,n,0), z(B,0)]
for fl in (16, 17):
    o=mgqp_b200.solve_quadprog_batched(*bare, flags=fl); torch.cuda.synchronize()
    t0=time.time(); o=mgqp_b200.solve_quadprog_batched(*bare, flags=fl); torch.cuda.synchronize(); dt=time.time()-t0
    fl_=(2*n**3/3+2*n*n)*B
    print('dense part flags',fl,': %.2f ms, %.3f TFLOP/s'%(dt*1e3, fl_/dt/1e12))
    if fl==16: xe=o.x.clone()
    else: print('  x0 rel diff', float(((o.x-xe).abs().amax(dim=1)/xe.abs().amax(dim=1).clamp(min=1)).max()))
P
