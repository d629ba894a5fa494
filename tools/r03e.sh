for v in nohyp nosd both; do
MGQP_B200_LIB=$PWD/variants/$v.so python - <<P
import sys, time, os; sys.path.insert(0,'.')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
dev=torch.device('cuda:0')
t=workloads.make_c4_device(592, dev)
ex=mgqp_b200.solve_quadprog_batched(*t, want_active_set=True)
fa=mgqp_b200.solve_quadprog_batched(*t, want_active_set=True, flags=1); torch.cuda.synchronize()
rel=((ex.x-fa.x).abs().amax(dim=1)/ex.x.abs().amax(dim=1).clamp(min=1.0))
print('$v','A equal',float((ex.A==fa.A).all(dim=1).double().mean()),'x rel max',float(rel.max()),'median',float(rel.median()),'iters eq',float((ex.iters==fa.iters).double().mean()))
P
done
