"""Small batches through every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgqp_b200
from mgqp_b200 import workloads
dev = torch.device("cuda:0")
def run(name, B, flags=0):
    b = workloads.make(name, B)
    t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in b.fields()]
    r = mgqp_b200.solve_quadprog_batched(*t, flags=flags, want_active_set=True)
    torch.cuda.synchronize()
    print(name, B, flags, r.status.bincount().tolist(), flush=True)
run("C2", 2 * 148 * 8 * 32 + 77)          # lane kernel, more than one round per warp, ragged last tile
run("C5", 4099, mgqp_b200.FLAG_RETRY_WITHOUT_CI)
run("C5D", 2051)
run("S0", 1500, mgqp_b200.FLAG_RETRY_WITHOUT_CI)
run("C3", 700)
b = workloads.make_random(40, 40, 2, 10, seed=3)   # CTA-per-QP kernel
t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in b.fields()]
for fl in (0, 1):
    r = mgqp_b200.solve_quadprog_batched(*t, flags=fl); torch.cuda.synchronize(); print("R40", fl, r.status.bincount().tolist(), flush=True)
# host pipeline: unit codes + expansion kernel, several chunks
b = workloads.make("C5", 9001)
r = mgqp_b200.solve_quadprog_batched_host(*b.fields(), chunk=2048)
print("host", np.bincount(r.status).tolist(), flush=True)
# cascade
s = workloads.make_scenarios(3000, seed=1)
st = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in s.items()}
c = mgqp_b200.solve_cascade_batched(st)
torch.cuda.synchronize()
print("cascade ok", flush=True)
