# is the wide-limits cascade really slower (r06h: 35 ms against 25.5 ms), or was that one stall inside three calls?
timeout 40 python - <<'P'
import sys; sys.path.insert(0,'.')
import numpy as np, torch, mgqp_b200
from mgqp_b200 import workloads
dev=torch.device('cuda:0')
cfg=workloads.WIDE_CONFIG
s=workloads.make_scenarios(400000, config=cfg, scale=0.01)
t={k: torch.from_numpy(v).to(dev) for k,v in s.items()}
def run(tag):
    ts=[]
    for _ in range(5):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); mgqp_b200.solve_cascade_batched(t, cfg); e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1),2))
    print(tag, ts, flush=True)
mgqp_b200.solve_cascade_batched(t, cfg); torch.cuda.synchronize()
run('fresh')
b=workloads.make('S0',100000)
x=[torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in b.fields()]
for _ in range(3): mgqp_b200.solve_quadprog_batched(*x)
torch.cuda.synchronize()
run('after classified S0 calls')
P
