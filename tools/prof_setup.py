"""Set-up only launch of the CTA-per-QP kernel (same G, g0 as C4, p = m = 0) for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgqp_b200
from mgqp_b200 import workloads
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
b = workloads.make("C4", B)
dev = "cuda:0"
G = torch.from_numpy(b.G).to(dev); g0 = torch.from_numpy(b.g0).to(dev)
z = lambda *s: torch.zeros(s, dtype=torch.float64, device=dev)
args = [G, g0, z(B, b.n, 0), z(B, 0), z(B, b.n, 0), z(B, 0)]
mgqp_b200.solve_quadprog_batched(*args, flags=16); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); mgqp_b200.solve_quadprog_batched(*args, flags=16); e1.record(); torch.cuda.synchronize()
print(f"setup only, B={B}: {e0.elapsed_time(e1):.2f} ms")
