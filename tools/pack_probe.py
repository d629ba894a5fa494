"""Host packer / packed-transport pipeline probe (run on the GPU box).  usage: pack_probe.py [B]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgqp_b200
from mgqp_b200 import workloads
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
b = workloads.make("C2", B)
pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in b.fields()]
L = mgqp_b200.lib()
n, m = b.n, b.m
val = torch.empty((B, m), dtype=torch.float64).pin_memory().numpy()
row = torch.empty((B, 48), dtype=torch.uint8).pin_memory().numpy()
CI = pinned[4]
for th in (1, 2, 4, 8, 16, 32):
    L.mgqp_pack_bound_columns(n, m, B, CI.ctypes.data, n * m, val.ctypes.data, m, row.ctypes.data, 48, th)
    t0 = time.perf_counter()
    L.mgqp_pack_bound_columns(n, m, B, CI.ctypes.data, n * m, val.ctypes.data, m, row.ctypes.data, 48, th)
    dt = time.perf_counter() - t0
    print(f"pack threads={th}: {B/dt:.3e} QP/s  ({B*n*m*8/dt/1e9:.1f} GB/s read)", flush=True)
res = mgqp_b200.QPResult(torch.empty((B, n), dtype=torch.float64).pin_memory().numpy(),
                         torch.empty((B,), dtype=torch.float64).pin_memory().numpy(),
                         torch.empty((B,), dtype=torch.int32).pin_memory().numpy())
for flags in (0, 32):
    for chunk in (0, 32768, 65536, 131072, 262144):
        mgqp_b200.solve_quadprog_batched_host(*pinned, chunk=chunk, out=res, flags=flags)
        t0 = time.perf_counter()
        for _ in range(3):
            mgqp_b200.solve_quadprog_batched_host(*pinned, chunk=chunk, out=res, flags=flags)
        dt = (time.perf_counter() - t0) / 3
        print(f"e2e flags={flags} chunk={chunk} threads={os.environ.get('MGQP_HOST_THREADS','auto')}: {B/dt:.3e} QP/s", flush=True)
