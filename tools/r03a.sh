python -m pytest tests/test_gpu_parity.py tests/test_cascade.py -m gpu -x -q 2>&1 | tail -3
python tools/cascade_probe.py 400000
python - <<'P'
import sys, time; sys.path.insert(0,'.')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
b=workloads.make('S0',400000); b.ci0[::3,3]=-7.0; b.ci0[::3,17]=-7.0
t=[torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in b.fields()]
for fl in (0,64):
    out=mgqp_b200.solve_quadprog_batched(*t, flags=fl); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): mgqp_b200.solve_quadprog_batched(*t, flags=fl, out=out)
    e1.record(); torch.cuda.synchronize()
    print('S0 (1/3 infeasible) flags',fl,'%.2f ms'%(e0.elapsed_time(e1)/3), np.unique(out.status.cpu().numpy(),return_counts=True))
P
