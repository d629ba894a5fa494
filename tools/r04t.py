import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgqp_b200
from mgqp_b200 import workloads
L = mgqp_b200.lib()
for name, B in (("C3", 100000), ("S0", 400000), ("S2", 400000)):
    b = workloads.make(name, B)
    pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in b.fields()]
    res = mgqp_b200.QPResult(torch.empty((B, b.n), dtype=torch.float64).pin_memory().numpy(),
                             torch.empty((B,), dtype=torch.float64).pin_memory().numpy(),
                             torch.empty((B,), dtype=torch.int32).pin_memory().numpy())
    per_qp = L.mgqp_algorithmic_bytes(b.n, b.p, b.m)
    for chunk in ((256 << 20) // per_qp & ~1023, 0, 8192):
        mgqp_b200.solve_quadprog_batched_host(*pinned, out=res, chunk=chunk)
        ts = []
        for _ in range(4):
            t0 = time.perf_counter()
            mgqp_b200.solve_quadprog_batched_host(*pinned, out=res, chunk=chunk)
            ts.append(time.perf_counter() - t0)
        up, down = mgqp_b200.last_transfer_bytes()
        print(f"e2e {name} chunk={chunk}: best {B/min(ts):.3e} median {B/sorted(ts)[2]:.3e} QP/s  h2d {up/B:.0f} B/QP", flush=True)
