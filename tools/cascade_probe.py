"""Throughput of the batched assembly + cascade (device-resident and from host state)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mgqp_b200
from mgqp_b200 import workloads
B = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
for name, cfg, scale in (("shipped", workloads.SHIPPED_CONFIG, 1.0), ("wide", workloads.WIDE_CONFIG, 0.01)):
    s = workloads.make_scenarios(B, config=cfg, scale=scale)
    t = {k: torch.from_numpy(v).cuda() for k, v in s.items()}
    r = mgqp_b200.solve_cascade_batched(t, cfg); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): mgqp_b200.solve_cascade_batched(t, cfg)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    sp = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in s.items()}
    mgqp_b200.solve_cascade_batched_host(sp, cfg)
    t0 = time.perf_counter(); h = mgqp_b200.solve_cascade_batched_host(sp, cfg); dt = time.perf_counter() - t0
    print(f"{name}: {B} scenarios, device {ms:.2f} ms = {B/ms*1e3:.3g} cascades/s; host {B/dt:.3g} cascades/s; "
          f"status0 {np.unique(r.status0.cpu().numpy(), return_counts=True)}; bytes {mgqp_b200.last_transfer_bytes()}")
