#!/usr/bin/env python
"""Benchmark of the mgqp hot path: batched QP solves/sec (fp64).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2|C5|C3|S0|S2]

A "step" is one pass of the solver over one batch of synthetic QPs.  At N=1 the workload is
BASELINE.json configs[1]: 1M independent random 7-DoF mgqp QPs (n=7, 1 equality, 42 bound
inequalities).  For N>1 (torchrun, one process per GPU) every rank holds its own 1M-QP shard
(different seed) -- the QPs are independent, so there is no collective on the data path and the
scaling is weak; `value` = all ranks' QPs / max-over-ranks device time.

Printed keys (one JSON line, rank 0):
  value     QPs/s with inputs already resident in HBM (CUDA events on the launching stream)
  e2e       QPs/s through the host-buffer plugin call solve_quadprog_batched_host(): pinned host
            inputs, H2D + solve + D2H all inside the timed region
  roofline  achieved algorithmic HBM bytes/s of the solver kernel vs the measured copy bandwidth
  cpu_baseline  the reference's shipped CPU solver (oracle/_ref) on a bounded sample, all host cores

--impl reference times the reference's own CPU implementation (oracle/_ref = the shipped
libquadprog.a; falls back to the oracle port where that was not built) on bounded samples of the
same workload; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

QPS_PER_GPU = {"C2": 1_000_000, "C5": 1_000_000, "C3": 100_000, "C4": 592, "S0": 400_000, "S2": 400_000}
CPU_SAMPLE = {"C2": 200_000, "C5": 200_000, "C3": 60_000, "C4": 16, "S0": 60_000, "S2": 100_000}


CONFIG_NOTE = {
    "C2": "BASELINE.json configs[1]: 1M independent random 7-DoF mgqp QPs",
    "C5": "BASELINE.json configs[4] shape: scenario sweep, ~10 % infeasible",
    "C3": "BASELINE.json configs[2]: task-space QPs with slacks",
    "C4": "BASELINE.json configs[3]: horizon-stacked MPC QPs, 7 DoF x 50 steps",
    "S0": "as-shipped mgqp level 0 (src/mgqp.cpp:1076-1143)",
    "S2": "as-shipped mgqp level 2",
}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            c = [v.strip() for v in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_rate(b, threads: int):
    """QPs/s of the reference's CPU solver on host batch b; returns (rate, kind)."""
    from oracle import oracle_py
    ref = oracle_py.reference()
    if ref is not None:
        t0 = time.perf_counter()
        ref.solve_batch(b.n, b.p, b.m, *b.fields(), threads=threads)
        return b.batch / (time.perf_counter() - t0), "reference"
    rest = oracle_py.restatement()
    t0 = time.perf_counter()
    rest.solve_batch(b.n, b.p, b.m, *b.fields(), threads=threads)
    return b.batch / (time.perf_counter() - t0), "port"


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    from mgqp_b200 import workloads
    name = args.workload
    nthreads = host_threads()
    sample = CPU_SAMPLE[name]
    b = workloads.make(name, sample)
    n, p, m = b.n, b.p, b.m
    rates = []
    kind = "reference"
    for i in range(args.warmup + args.steps):
        r, kind = cpu_reference_rate(b, nthreads)
        if i >= args.warmup:
            rates.append(r)
    ms = 1e3 * sample * np.mean([1.0 / r for r in rates])
    value = sample / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "batched QP solves/sec (fp64)", "value": value, "unit": "QP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{name}: n={n} p={p} m={m}; CPU step = {sample} QPs (bounded sample of the "
                               f"{QPS_PER_GPU[name]}-QP workload)", "host_threads": nthreads},
        "cpu_baseline": {"value": value, "unit": "QP/s", "cores": nthreads, "kind": kind,
                         "sample": f"{sample} {name} QPs per step, all host threads; "
                                   + ("shipped libquadprog.a (GCC 5.4 -O0) via oracle/_ref" if kind == "reference"
                                      else "oracle/gi_oracle.c -O2 (oracle/_ref not built)")},
        "e2e": {"value": value, "unit": "QP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    # one rank per GPU shares the box's host cores: the host packer of the e2e pipeline gets its share
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    os.environ.setdefault("MGQP_HOST_THREADS", str(max(1, host_threads() // max(1, local_world))))
    import mgqp_b200
    from mgqp_b200 import workloads

    assert torch.cuda.is_available(), "bench.py (impl=ours) needs a GPU; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    name = args.workload
    B = args.batch or QPS_PER_GPU[name]
    from mgqp_b200 import sharding
    b = workloads.make(name, B, seed=sharding.rank_seed(1234 + workloads._CONFIG_ID[name], rank) if rank else None)
    n, p, m = b.n, b.p, b.m
    bytes_per_qp = mgqp_b200.algorithmic_bytes(n, p, m)

    # ---- device-resident leg
    host_pinned = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in b.fields()]
    dev_in = [t.to(dev, non_blocking=True) for t in host_pinned]
    torch.cuda.synchronize()
    out = mgqp_b200.solve_quadprog_batched(*dev_in, flags=args.flags)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        mgqp_b200.solve_quadprog_batched(*dev_in, flags=args.flags, out=out)
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    l0 = mgqp_b200.kernel_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        mgqp_b200.solve_quadprog_batched(*dev_in, flags=args.flags, out=out)
        ev[i + 1].record()
    barrier()
    launches = mgqp_b200.kernel_launch_count() - l0
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * B * args.steps / (total_ms / 1e3)
    status = out.status.cpu().numpy()

    # ---- end-to-end leg: host (pinned) buffers through the plugin call, copies inside the timing
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    host_np = [t.numpy() for t in host_pinned]
    res = mgqp_b200.QPResult(torch.empty((B, n), dtype=torch.float64).pin_memory().numpy(),
                             torch.empty((B,), dtype=torch.float64).pin_memory().numpy(),
                             torch.empty((B,), dtype=torch.int32).pin_memory().numpy())
    mgqp_b200.solve_quadprog_batched_host(*host_np, device=local_rank, flags=args.flags, out=res)  # warm
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mgqp_b200.solve_quadprog_batched_host(*host_np, device=local_rank, flags=args.flags, out=res)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()  # sampled over both timed regions (device-resident leg and e2e leg)
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(t.item())
    # bytes that actually crossed the link in the last call (the library counts its own copies): where a
    # chunk's CI is a set of bounds it travels column-compressed (m*9 bytes, mgqp_batch::CI_val) instead
    # of as the dense n*m*8-byte block the caller handed over
    h2d, d2h = mgqp_b200.last_transfer_bytes()
    host_bytes = sum(a.nbytes for a in host_np)
    e2e_same = bool(np.array_equal(res.status, status) and np.array_equal(res.x, out.x.cpu().numpy()))

    if rank != 0:
        return
    # ---- parity spot check + CPU baseline on a bounded sample of the same inputs (rank 0, N=1 only)
    line_cpu = None
    parity = None
    if world == 1 and not args.no_cpu:
        from oracle import oracle_py
        ns = min(B, CPU_SAMPLE[name])
        sb = b.slice(0, ns)
        nthreads = host_threads()
        rate, kind = cpu_reference_rate(sb, nthreads)
        line_cpu = {"value": rate, "unit": "QP/s", "cores": nthreads, "kind": kind,
                    "sample": f"first {ns} QPs of the step's batch, {nthreads} host threads, "
                              + ("shipped libquadprog.a (GCC 5.4 -O0) via oracle/_ref" if kind == "reference"
                                 else "oracle/gi_oracle.c -O2")}
        chk = oracle_py.restatement().solve_batch(n, p, m, *sb.fields(), max_iter=100 + 10 * (p + m), threads=nthreads)
        xg = out.x[:ns].cpu().numpy()
        fg = out.f[:ns].cpu().numpy()
        parity = {"sample": ns, "status_equal": bool(np.array_equal(status[:ns], chk["status"])),
                  "x_bit_identical": bool(np.array_equal(xg, chk["x"])),
                  "f_bit_identical": bool(np.array_equal(fg, chk["f"], equal_nan=True))}
    peak, peak_src = peaks()
    kern_ms = float(np.mean(step_ms))  # one launch per step: the step IS the kernel (+ a 4-byte memset)
    gw, warps, smem = mgqp_b200.plan(n, p, m)
    roof_note = "every input once, x/f/status once (SURVEY.md 8d)"
    if gw == 0 and not (args.flags & 8):
        # CTA-per-QP shapes: CI (8*n*m bytes) no longer fits on chip and the algorithm itself reads all of
        # it in every violated-constraint scan, so that is the algorithmic traffic
        probe = mgqp_b200.solve_quadprog_batched(*dev_in, flags=args.flags, want_active_set=True)
        torch.cuda.synchronize()
        scans = float(probe.iters.double().mean().item()) + 1.0
        bytes_per_qp = int(bytes_per_qp + (scans - 1.0) * 8 * n * m)
        roof_note = f"inputs once + CI re-read in each of the {scans:.1f} scans per QP (SURVEY.md 8a a7)"
    achieved = B * bytes_per_qp / (kern_ms / 1e3) / 1e9
    kname = {1: "gi_lane_kernel<%d> (one QP per lane)" % n, 0: "gi_large_kernel (one QP per CTA)"}.get(
        gw, "gi_small_kernel<%d> (%d lanes per QP)" % (gw, gw))
    if args.flags & 8:
        kname = "gi_small_kernel (lane groups, forced)"
    if args.flags & 16:
        kname = "gi_large_kernel (one QP per CTA, forced)"
    line = {
        "metric": "batched QP solves/sec (fp64)", "value": value, "unit": "QP/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{name}: {B} QPs per GPU, n={n} p={p} m={m} ({CONFIG_NOTE[name]})",
                   "l2": f"inputs {B * bytes_per_qp / 1e6:.0f} MB per step > 126 MB L2, read once per step",
                   "mode": "fast_math" if args.flags & 1 else "exact (bit-identical to QuadProg++)",
                   "kernel": f"{kname}, {warps} warps/CTA, {smem} B smem/CTA", "flags": args.flags,
                   "status_counts": np.bincount(status, minlength=5).tolist()},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "QP/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "matches_device_leg": e2e_same, "host_input_bytes_per_step": int(host_bytes),
                "api": "solve_quadprog_batched_host (C ABI, pinned host buffers, chunked 2-stream pipeline; CI is "
                       "packed on the host cores into (row, value) columns when every column is a bound"
                       + (" -- disabled by --flags 32)" if args.flags & 32 else ")")},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "bytes_per_qp": bytes_per_qp, "bytes_note": roof_note,
                     "kernel_ms": kern_ms},
    }
    if gw == 0 and not (args.flags & 8):
        # SURVEY.md 8d for the horizon-stacked shape: the dense factorisation part on its own
        # (Cholesky n^3/3 + triangular inverse n^3/3 + solve 2n^2 flops), timed by solving the same G, g0
        # with no constraints, against an fp64 GEMM peak measured here with torch.matmul.
        z = lambda *sh: torch.zeros(sh, dtype=torch.float64, device=dev)
        bare = [dev_in[0], dev_in[1], z(B, n, 0), z(B, 0), z(B, n, 0), z(B, 0)]
        mgqp_b200.solve_quadprog_batched(*bare, flags=args.flags | 16)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            mgqp_b200.solve_quadprog_batched(*bare, flags=args.flags | 16)
        e1.record()
        torch.cuda.synchronize()
        dense_ms = e0.elapsed_time(e1) / 3
        flops = (2.0 * n ** 3 / 3.0 + 2.0 * n * n) * B
        a = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
        torch.matmul(a, a)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            torch.matmul(a, a)
        e1.record()
        torch.cuda.synchronize()
        p64 = 3 * 2.0 * 4096 ** 3 / (e0.elapsed_time(e1) / 1e3) / 1e12
        line["dense_part"] = {"what": "set-up only (Cholesky, J = L^-T, x0): same G, g0, p = m = 0", "ms": dense_ms,
                              "tflops": flops / (dense_ms / 1e3) / 1e12, "peak_tflops": p64,
                              "peak_source": "torch.matmul fp64 4096^3 on this GPU",
                              "frac": flops / (dense_ms / 1e3) / 1e12 / p64,
                              "note": "plain FP64 FMA units in reference summation order; no DMMA (it would change "
                                      "the rounding order of the exact mode)"}
    if line_cpu:
        line["cpu_baseline"] = line_cpu
    if parity:
        line["parity_check"] = parity
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            line["roofline"]["traffic"] = json.load(open(tr)).get(name)
        except Exception:
            pass
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(QPS_PER_GPU))  # C2 = the headline config
    ap.add_argument("--batch", type=int, default=0, help="QPs per GPU per step (default: the workload's)")
    ap.add_argument("--flags", type=int, default=0, help="MGQP_FLAG_* bits (1 = fast math)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
