"""One or a few launches of the solver on a seeded batch (for ncu / quick timing)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import mgqp_b200
from mgqp_b200 import workloads

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="C2")
ap.add_argument("--batch", type=int, default=200000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--compressed", action="store_true", help="hand CI over as (value, row) columns")
a = ap.parse_args()
b = workloads.make(a.workload, a.batch)
t = [torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in b.fields()]
kw = {}
if a.compressed:
    val, row = mgqp_b200.pack_bound_columns(b.CI)
    kw = dict(CI_val=torch.from_numpy(val).cuda(), CI_row=torch.from_numpy(row).cuda())
    t[4] = None
out = mgqp_b200.solve_quadprog_batched(*t, flags=a.flags, want_active_set=True, **kw)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(a.reps):
    mgqp_b200.solve_quadprog_batched(*t, flags=a.flags, out=out, **kw)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / a.reps
it = out.iters.cpu().numpy()
print(f"{a.workload} B={a.batch} flags={a.flags}: {ms:.3f} ms/launch, {a.batch / ms * 1e3:.4g} QP/s; "
      f"steps mean {it.mean():.2f} max {it.max()}; statuses {np.bincount(out.status.cpu().numpy()).tolist()}")
