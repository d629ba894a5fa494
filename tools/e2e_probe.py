import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgqp_b200
from mgqp_b200 import workloads
B = 400000
b = workloads.make("C2", B)
pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in b.fields()]
for flags in (0, 8):
    for chunk in (0, 20000, 200000):
        mgqp_b200.solve_quadprog_batched_host(*pinned, flags=flags, chunk=chunk)
        t0 = time.perf_counter()
        mgqp_b200.solve_quadprog_batched_host(*pinned, flags=flags, chunk=chunk)
        dt = time.perf_counter() - t0
        print(f"flags={flags} chunk={chunk}: {dt*1e3:.1f} ms  {B/dt:.3e} QP/s")
# device-resident small batches
t = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in b.fields()]
for nb in (20000, 81920, 400000):
    tt = [x[:nb] for x in t]
    for flags in (0, 8):
        out = mgqp_b200.solve_quadprog_batched(*tt, flags=flags)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            mgqp_b200.solve_quadprog_batched(*tt, flags=flags, out=out)
        e1.record(); torch.cuda.synchronize()
        print(f"device nb={nb} flags={flags}: {e0.elapsed_time(e1)/3:.3f} ms")
