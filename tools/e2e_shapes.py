import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgqp_b200
from mgqp_b200 import workloads
for name, B in (("S0", 400000), ("S2", 400000), ("C5", 1000000)):
    b = workloads.make(name, B)
    pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in b.fields()]
    res = mgqp_b200.QPResult(torch.empty((B, b.n), dtype=torch.float64).pin_memory().numpy(),
                             torch.empty((B,), dtype=torch.float64).pin_memory().numpy(),
                             torch.empty((B,), dtype=torch.int32).pin_memory().numpy())
    for flags in (0, 32, 64):
        mgqp_b200.solve_quadprog_batched_host(*pinned, out=res, flags=flags)
        t0 = time.perf_counter()
        for _ in range(3):
            mgqp_b200.solve_quadprog_batched_host(*pinned, out=res, flags=flags)
        dt = (time.perf_counter() - t0) / 3
        up, down = mgqp_b200.last_transfer_bytes()
        print(f"e2e {name} flags={flags}: {B/dt:.3e} QP/s  h2d {up/B:.0f} B/QP  statuses {np.bincount(res.status).tolist()}", flush=True)
