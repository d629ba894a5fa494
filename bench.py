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

QPS_PER_GPU = {"C2": 1_000_000, "C5": 1_000_000, "C5D": 1_000_000, "C3": 100_000, "C4": 592, "S0": 400_000,
               "S2": 400_000}
# CPU baseline samples: (QPs solved on all host threads, QPs solved on one thread) -- about a second or two each
CPU_SAMPLE = {"C2": (200_000, 20_000), "C5": (200_000, 20_000), "C5D": (200_000, 20_000), "C3": (60_000, 6_000),
              "C4": (16, 1), "S0": (60_000, 6_000), "S2": (100_000, 10_000)}
# the sub-records bench.py appends to the headline line at N = 1: (workload, QPs per step, steps)
EXTRA_CONFIGS = [("C5", 1_000_000, 5), ("C5D", 1_000_000, 5), ("C3", 100_000, 5), ("S0", 400_000, 5),
                 ("S2", 400_000, 5)]
C4_FULL = 10_000          # BASELINE.json configs[3]: 10k horizon-stacked QPs (generated on the device)
C5_SHARDED_TOTAL = 10_000_000  # BASELINE.json configs[4]: 10M-QP sweep, strong-scaled over the ranks


CONFIG_NOTE = {
    "C2": "BASELINE.json configs[1]: 1M independent random 7-DoF mgqp QPs",
    "C5": "BASELINE.json configs[4] shape: scenario sweep, ~10 % infeasible",
    "C5D": "configs[4] with a second equality: ~10 % infeasible, ~1 % linearly dependent equalities (SURVEY.md 8d)",
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

    def __init__(self, index: int, period_ms: int = 20):
        self.index = index
        self.period_ms = period_ms
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", str(self.period_ms)], stdout=subprocess.PIPE,
                                         text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_ready(self, timeout: float = 5.0):
        """Block until nvidia-smi has delivered its first sample (its start-up can outlast a 20 ms timed region)."""
        t0 = time.perf_counter()
        while self.proc and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def mark(self) -> int:
        """Index of the next sample: pass it to stop() to keep only the samples taken from here on."""
        return len(self.lines)

    def stop(self, first: int = 0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(1.5e-3 * self.period_ms)  # let the sample in flight arrive; longer would only add idle samples
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = self.lines[first:] or self.lines  # (the region was shorter than one period: the nearest samples)
        for ln in lines:
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


def cpu_contenders(b, name: str):
    """both CPU contenders -- (i) the shipped libquadprog.a (GCC 5.4 -O0, oracle/_ref) and
    (ii) the -O2 restatement (oracle/gi_oracle.c) -- each on all host threads and on one thread, on bounded
    samples of batch b.  QPs/s; None where oracle/_ref was not built."""
    from oracle import oracle_py
    nthreads = host_threads()
    n_all, n_one = CPU_SAMPLE[name]
    n_all, n_one = min(n_all, b.batch), min(n_one, b.batch)
    out = {"cores": nthreads, "sample_all_threads": n_all, "sample_one_thread": n_one, "unit": "QP/s"}
    ref, rest = oracle_py.reference(), oracle_py.restatement()
    for key, solver in (("shipped_O0", ref), ("restatement_O2", rest)):
        if solver is None:
            out[key] = None
            continue
        rec = {}
        for label, cnt, thr in (("all_threads", n_all, nthreads), ("one_thread", n_one, 1)):
            sb = b.slice(0, cnt)
            t0 = time.perf_counter()
            solver.solve_batch(sb.n, sb.p, sb.m, *sb.fields(), threads=thr)
            rec[label] = cnt / (time.perf_counter() - t0)
        out[key] = rec
    return out


def parity_vs_oracle(b, got_x, got_f, got_status, ns):
    from oracle import oracle_py
    sb = b.slice(0, ns)
    chk = oracle_py.restatement().solve_batch(sb.n, sb.p, sb.m, *sb.fields(), max_iter=100 + 10 * (sb.p + sb.m),
                                              threads=host_threads())
    return {"sample": int(ns), "status_equal": bool(np.array_equal(got_status[:ns], chk["status"])),
            "x_bit_identical": bool(np.array_equal(got_x[:ns], chk["x"], equal_nan=True)),
            "f_bit_identical": bool(np.array_equal(got_f[:ns], chk["f"], equal_nan=True))}


def timed_launches(mgqp_b200, dev_in, steps, warmup, flags, out=None, **kw):
    """(mean ms per launch, result) of the device-resident solver, CUDA events on torch's current stream."""
    import torch
    out = mgqp_b200.solve_quadprog_batched(*dev_in, flags=flags, out=out, **kw)
    for _ in range(max(0, warmup - 1)):
        mgqp_b200.solve_quadprog_batched(*dev_in, flags=flags, out=out, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        mgqp_b200.solve_quadprog_batched(*dev_in, flags=flags, out=out, **kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def kernel_name(mgqp_b200, n, p, m):
    gw, warps, smem = mgqp_b200.plan(n, p, m)
    k = {1: "gi_lane_kernel<%d> (one QP per lane)" % n, 0: "gi_large_kernel (one QP per CTA)"}.get(
        gw, "gi_small_kernel<%d> (%d lanes per QP)" % (gw, gw))
    return f"{k}, {warps} warps/CTA, {smem} B smem/CTA"


def extra_config_record(mgqp_b200, name, B, steps, dev, peak):
    """One sub-record: a BASELINE config other than the headline, device-resident, with parity and CPU baselines."""
    import torch
    from mgqp_b200 import workloads
    b = workloads.make(name, B)
    n, p, m = b.n, b.p, b.m
    dev_in = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in b.fields()]
    ms, out = timed_launches(mgqp_b200, dev_in, steps, 3, 0)
    status = out.status.cpu().numpy()
    ns = min(B, CPU_SAMPLE[name][0])
    bytes_per_qp = mgqp_b200.algorithmic_bytes(n, p, m)
    achieved = B * bytes_per_qp / (ms / 1e3) / 1e9
    # the same batch with its inequality block in column-compressed form (mgqp_batch::CI_val / CI_row: what the host
    # pipeline and the cascade hand the kernel): the group kernel keeps it compressed in shared memory, more warps fit
    comp = None
    if n > 8:
        try:
            packed = mgqp_b200.pack_bound_columns(b.CI)
            if packed is not None:
                v = torch.from_numpy(packed[0]).to(dev)
                r = torch.from_numpy(packed[1]).to(dev)
                kw = dict(CI_val=v, CI_row=r)
                oc = mgqp_b200.solve_quadprog_batched(dev_in[0], dev_in[1], dev_in[2], dev_in[3], None, dev_in[5], **kw)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(steps):
                    mgqp_b200.solve_quadprog_batched(dev_in[0], dev_in[1], dev_in[2], dev_in[3], None, dev_in[5], out=oc, **kw)
                e1.record()
                torch.cuda.synchronize()
                cms = e0.elapsed_time(e1) / steps
                comp = {"value": B / (cms / 1e3), "unit": "QP/s", "ms_per_step": cms,
                        "same_answers": bool(torch.equal(oc.x, out.x) and torch.equal(oc.status, out.status)),
                        "what": "device-resident, CI as (value, row) columns instead of the dense block"}
                del oc, v, r
        except Exception as e:  # never let the extra measurement cost the record
            comp = {"error": repr(e)}
    host_np = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in b.fields()]
    # pinned result buffers, as in the headline's e2e leg (pageable ones turn the downloads into staged, blocking
    # copies: 8.7e6 instead of 1.8e7 QP/s on S0)
    res = mgqp_b200.QPResult(torch.empty((B, n), dtype=torch.float64).pin_memory().numpy(),
                             torch.empty((B,), dtype=torch.float64).pin_memory().numpy(),
                             torch.empty((B,), dtype=torch.int32).pin_memory().numpy())
    mgqp_b200.solve_quadprog_batched_host(*host_np, device=dev.index or 0, out=res)  # warm (staging arenas)
    e2e_steps = 3
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mgqp_b200.solve_quadprog_batched_host(*host_np, device=dev.index or 0, out=res)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    h2d, d2h = mgqp_b200.last_transfer_bytes()
    rec = {"workload": f"{name}: {B} QPs, n={n} p={p} m={m} ({CONFIG_NOTE[name]})", "value": B / (ms / 1e3),
           "unit": "QP/s", "ms_per_step": ms, "steps": steps, "kernel": kernel_name(mgqp_b200, n, p, m),
           "status_counts": np.bincount(status & 15, minlength=5).tolist(),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "bytes_per_qp": bytes_per_qp},
           "e2e": {"value": B / e2e_s, "unit": "QP/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                   "steps": e2e_steps, "note": "solve_quadprog_batched_host on pinned host arrays (inputs and results), after one warm call",
                   "matches_device_leg": bool(np.array_equal(res.status, status))},
           "parity_check": parity_vs_oracle(b, out.x.cpu().numpy(), out.f.cpu().numpy(), status, ns),
           "cpu_baseline": cpu_contenders(b, name)}
    if comp is not None:
        rec["compressed_input"] = comp
    del dev_in, out
    torch.cuda.empty_cache()
    return rec


def c4_record(mgqp_b200, dev, peak, B):
    """BASELINE.json configs[3] at its stated size: B horizon-stacked QPs (n = 350, m = 2100) generated on the
    device (69 GB at 10k QPs).  The HBM fraction uses SURVEY.md 8(d)'s one-read bytes; the CI re-reads the
    algorithm itself performs (one pass over CI per violated-constraint scan) are reported next to it, never
    folded into it."""
    import torch
    from mgqp_b200 import workloads
    n, p, m = workloads.SHAPES["C4"]
    free, _ = torch.cuda.mem_get_info(dev)
    per_qp = 8 * (n * n + n + n * m + m)
    B = int(min(B, (free * 0.8 - (6 << 30)) // (per_qp + 8 * n * n)))  # G is built from an n*n temporary
    t = workloads.make_c4_device(B, dev)
    torch.cuda.synchronize()
    ms, out = timed_launches(mgqp_b200, t, 1, 1, 0, want_active_set=True)
    status = out.status.cpu().numpy()
    iters = out.iters.double()
    ns = 8
    sample = workloads.QPBatch("C4", n, p, m, *[a[:ns].cpu().numpy() for a in t])
    bytes_per_qp = mgqp_b200.algorithmic_bytes(n, p, m)
    achieved = B * bytes_per_qp / (ms / 1e3) / 1e9
    scans = float(iters.mean().item()) + 1.0
    x = out.x
    # size-independent property at full size: every OK answer is primal feasible to 1e-9 (relative to the bound)
    CI, ci0 = t[4], t[5]
    # (the solver's own stopping rule: total violation -sum min(0, s_i) <= m * eps * trace(G) * trace(J) * 100 with
    # trace(J) <= n for G >= I)
    viol = 0
    for lo in range(0, B, 256):
        hi = min(B, lo + 256)
        s = torch.einsum("bnm,bn->bm", CI[lo:hi], x[lo:hi]) + ci0[lo:hi]
        ok = (out.status[lo:hi] == 0)
        thr = 100.0 * m * 2.220446049250313e-16 * torch.diagonal(t[0][lo:hi], dim1=1, dim2=2).sum(dim=1) * n
        viol += int((((-s.clamp(max=0.0).sum(dim=1)) > thr) & ok).sum().item())
    rec = {"workload": f"C4: {B} QPs, n={n} p={p} m={m} ({CONFIG_NOTE['C4']}); generated on the device, "
                       f"{B * per_qp / 1e9:.1f} GB resident", "value": B / (ms / 1e3), "unit": "QP/s",
           "ms_per_step": ms, "steps": 1, "kernel": kernel_name(mgqp_b200, n, p, m),
           "status_counts": np.bincount(status & 15, minlength=5).tolist(),
           "mean_steps_per_qp": float(iters.mean().item()),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "bytes_per_qp": bytes_per_qp,
                        "ci_reread_bytes_per_qp": int((scans - 1.0) * 8 * n * m),
                        "note": "frac = SURVEY.md 8(d) one-read bytes / time; the algorithm re-reads CI (8nm bytes) in "
                                "each of its scans: that traffic is reported separately, not as useful work"},
           "primal_infeasible_answers": viol,
           "parity_check": parity_vs_oracle(sample, x[:ns].cpu().numpy(), out.f[:ns].cpu().numpy(), status, ns),
           "cpu_baseline": cpu_contenders(sample, "C4")}
    # SURVEY.md 8d for the horizon-stacked shape: the dense factorisation part on its own (Cholesky n^3/3 +
    # triangular inverse n^3/3 + solve 2n^2 flops): same G, g0 with no constraints, against an fp64 GEMM peak
    # measured here with torch.matmul.
    # the tensor-core mode (MGQP_FLAG_FAST_MATH: blocked Cholesky / triangular inverse on fp64 DMMA, FMA elsewhere):
    # held to the north-star gate -- identical verdicts and active sets, x and f within 1e-9 relative
    ms_fast, out_fast = timed_launches(mgqp_b200, t, 1, 1, 1, want_active_set=True)
    relx = ((out_fast.x - x).abs().amax(dim=1) / x.abs().amax(dim=1).clamp(min=1.0)).max().item()
    relf = ((out_fast.f - out.f).abs() / out.f.abs().clamp(min=1.0)).max().item()
    rec["fast_math_dmma"] = {"value": B / (ms_fast / 1e3), "unit": "QP/s", "ms_per_step": ms_fast,
                             "gate": {"status_equal": bool(torch.equal(out_fast.status, out.status)),
                                      "active_sets_equal": bool(torch.equal(out_fast.A, out.A)),
                                      "x_rel_max": relx, "f_rel_max": relf, "within_1e-9": bool(relx <= 1e-9 and relf <= 1e-9)}}
    del out_fast
    # whole waves of the resident CTAs (four per SM for this shape, gi_large_kernel.cuh), so that the set-up's rate is
    # not quantised
    Bd = min(B, 4 * 4 * torch.cuda.get_device_properties(dev).multi_processor_count)
    z = lambda *sh: torch.zeros(sh, dtype=torch.float64, device=dev)
    bare = [t[0][:Bd], t[1][:Bd], z(Bd, n, 0), z(Bd, 0), z(Bd, n, 0), z(Bd, 0)]
    dense_ms, _ = timed_launches(mgqp_b200, bare, 2, 1, 16)
    dense_ms_dmma, _ = timed_launches(mgqp_b200, bare, 2, 1, 17)
    flops = (2.0 * n ** 3 / 3.0 + 2.0 * n * n) * Bd
    a = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
    torch.matmul(a, a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        torch.matmul(a, a)
    e1.record()
    torch.cuda.synchronize()
    p64 = 3 * 2.0 * 4096 ** 3 / (e0.elapsed_time(e1) / 1e3) / 1e12
    rec["dense_part"] = {"what": f"set-up only (Cholesky, J = L^-T, x0) of {Bd} QPs: same G, g0, p = m = 0",
                         "ms": dense_ms, "tflops": flops / (dense_ms / 1e3) / 1e12, "peak_tflops": p64,
                         "peak_source": "torch.matmul fp64 4096^3 on this GPU",
                         "frac": flops / (dense_ms / 1e3) / 1e12 / p64,
                         "note": "exact mode: the reference's summation order on the FP64 FMA units",
                         "dmma": {"what": "the same under MGQP_FLAG_FAST_MATH: blocked Cholesky + triangular inverse on "
                                          "mma.sync.m8n8k4.f64 (SASS DMMA)", "ms": dense_ms_dmma,
                                  "tflops": flops / (dense_ms_dmma / 1e3) / 1e12,
                                  "frac": flops / (dense_ms_dmma / 1e3) / 1e12 / p64}}
    del t, out, bare
    torch.cuda.empty_cache()
    return rec


def c1_latency_record(mgqp_b200):
    """BASELINE.json configs[0]: ONE QP per control cycle through the C++ drop-in solve_quadprog() (the symbol
    src/mgqp.cpp:708 links against), as-shipped shapes S0 and S2; median wall-clock microseconds per call next to
    the CPU solvers' per-QP time on one thread."""
    import ctypes as C
    from mgqp_b200 import workloads
    from oracle import oracle_py
    L = mgqp_b200.lib()
    dp = C.POINTER(C.c_double)
    L.mgqp_dropin_latency.restype = C.c_int
    L.mgqp_dropin_latency.argtypes = [C.c_int32] * 3 + [dp] * 6 + [C.c_int, dp, dp, dp]
    rec = {}
    for name in ("S0", "S2"):
        b = workloads.make(name, 256)
        reps = 300
        us = np.zeros(reps)
        xs = np.zeros(b.n)
        f = C.c_double(0.0)
        a = [np.ascontiguousarray(v[3]) for v in b.fields()]  # QP 3 of the batch
        rc = L.mgqp_dropin_latency(b.n, b.p, b.m, *[v.ctypes.data_as(dp) for v in a], reps, us.ctypes.data_as(dp),
                                   xs.ctypes.data_as(dp), C.byref(f))
        one = oracle_py.restatement().solve_one(*a)
        cpu = {}
        ref, rest = oracle_py.reference(), oracle_py.restatement()
        for key, solver in (("shipped_O0_us_per_qp", ref), ("restatement_O2_us_per_qp", rest)):
            if solver is None:
                cpu[key] = None
                continue
            t0 = time.perf_counter()
            solver.solve_batch(b.n, b.p, b.m, *b.fields(), threads=1)
            cpu[key] = 1e6 * (time.perf_counter() - t0) / b.batch
        rec[name] = {"shape": f"n={b.n} p={b.p} m={b.m}", "calls": reps, "rc": int(rc),
                     "median_us": float(np.median(us[20:])), "p10_us": float(np.percentile(us[20:], 10)),
                     "p90_us": float(np.percentile(us[20:], 90)), "first_call_us": float(us[0]),
                     "bit_identical_to_oracle": bool(np.array_equal(xs, one["x"]) and f.value == one["f"]),
                     "cpu_one_thread": cpu}
    rec["what"] = ("C++ drop-in solve_quadprog(ArrayHH...) at B = 1: host copy into pinned staging, one H2D copy, one "
                   "kernel launch, one D2H copy, one stream synchronisation per call")
    return rec


def cascade_record(mgqp_b200, dev):
    """the solver's two callers batched -- on-device QP assembly (updateHook) and the priority
    cascade (solveNextHierarchy) of the shipped 7-DoF wiring, from ~0.5 KB of robot state per scenario.  One
    cascade = two levels, each a solve plus (where the level is infeasible) solveNextStep's retry."""
    import torch
    from mgqp_b200 import workloads
    from oracle import cascade_oracle as co, oracle_py
    rec = {}
    B = 400_000
    for name, cfg, scale in (("shipped", workloads.SHIPPED_CONFIG, 1.0), ("wide_limits", workloads.WIDE_CONFIG, 0.01)):
        s = workloads.make_scenarios(B, config=cfg, scale=scale)
        t = {k: torch.from_numpy(v).to(dev) for k, v in s.items()}
        r = mgqp_b200.solve_cascade_batched(t, cfg, taps=True)
        # warm calls of the timed form: the call above (with taps) leaves the library's memory pool in another state, and
        # the first plain calls after it grow it again -- r06h caught 37 ms of that inside three timed calls of 22.7 ms
        # (`tools/r06i.sh`: five isolated calls 22.79 / 22.69 / 22.68 / 22.69 / 22.67 ms)
        for _ in range(2):
            mgqp_b200.solve_cascade_batched(t, cfg)
        torch.cuda.synchronize()
        l0 = mgqp_b200.kernel_launch_count()
        calls = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(calls):
            mgqp_b200.solve_cascade_batched(t, cfg)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / calls
        launches = (mgqp_b200.kernel_launch_count() - l0) // calls
        sp = {k: torch.from_numpy(v).pin_memory().numpy() for k, v in s.items()}
        mgqp_b200.solve_cascade_batched_host(sp, cfg, device=dev.index or 0)
        t0 = time.perf_counter()
        h = mgqp_b200.solve_cascade_batched_host(sp, cfg, device=dev.index or 0)
        e2e_s = time.perf_counter() - t0
        up, down = mgqp_b200.last_transfer_bytes()
        # parity on a sample: the restatement fed with the device's own projector / barrier limits must agree bit for bit
        ns = 4000
        sub = {k: v[:ns] for k, v in s.items()}
        exp = co.cascade(oracle_py.restatement(), sub, cfg, threads=host_threads(), Z_override=r.Z[:ns].cpu().numpy(),
                         limits_override=r.ci0[:ns].cpu().numpy().astype(np.float32))
        t0 = time.perf_counter()
        co.cascade(oracle_py.restatement(), sub, cfg, threads=host_threads())
        cpu_s = time.perf_counter() - t0
        st0 = r.status0.cpu().numpy()
        rec[name] = {"scenarios": B, "value": B / (ms / 1e3), "unit": "cascades/s", "ms_per_step": ms,
                     "kernel_launches_per_step": int(launches),
                     "level0_retried_frac": float(((st0 & 16) != 0).mean()),
                     "e2e": {"value": B / e2e_s, "unit": "cascades/s", "h2d_bytes_per_step": int(up),
                             "d2h_bytes_per_step": int(down), "matches_device_leg": bool(np.array_equal(
                                 h.tracking, r.tracking.cpu().numpy()))},
                     "parity_check": {"sample": ns, "tracking_bit_identical": bool(np.array_equal(
                         r.tracking[:ns].cpu().numpy(), exp["tracking"])),
                         "torques_bit_identical": bool(np.array_equal(r.torques[:ns].cpu().numpy(), exp["torques"])),
                         "status_equal": bool(np.array_equal(st0[:ns], exp["status0"])),
                         "oracle": "oracle/cascade_oracle.py fed with the device's projector Z (parity unpinned "
                                   "against Eigen: see that file)"},
                     "cpu_baseline": {"value": ns / cpu_s, "unit": "cascades/s", "cores": host_threads(), "kind": "port",
                                      "sample": f"{ns} scenarios: numpy float32 assembly + LAPACK SVD + the -O2 solver "
                                                "restatement on all host threads"}}
        del t, r
        torch.cuda.empty_cache()
    return rec


def c5_sharded_record(mgqp_b200, dev, rank, world, dist):
    """BASELINE.json configs[4]: a 10M-QP sweep with mixed feasible / infeasible instances, STRONG-scaled: the same
    10M QPs are split by QP index over the ranks (one GPU each, no data-path collective).  Every rank's shard is a
    1M-QP seeded batch repeated on the device; time = max over ranks."""
    import torch
    from mgqp_b200 import sharding, workloads
    lo, hi = sharding.shard_range(C5_SHARDED_TOTAL, rank, world)
    mine = hi - lo
    base = workloads.make("C5", min(1_000_000, mine), seed=sharding.rank_seed(1234 + 5, rank))
    reps = -(-mine // base.batch)
    dev_in = [torch.from_numpy(np.ascontiguousarray(a)).to(dev).repeat((reps,) + (1,) * (a.ndim - 1))[:mine]
              for a in base.fields()]
    out = mgqp_b200.solve_quadprog_batched(*dev_in)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mgqp_b200.solve_quadprog_batched(*dev_in, out=out)
    e1.record()
    torch.cuda.synchronize()
    my_ms = e0.elapsed_time(e1)
    ms = torch.tensor([my_ms], dtype=torch.float64, device=dev)
    counts = torch.bincount(out.status.long() & 15, minlength=5).double()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts)
    value = sharding.whole_job_rate(mine, my_ms / 1e3)  # all ranks' QPs / max-over-ranks time
    del dev_in, out
    torch.cuda.empty_cache()
    return {"workload": f"C5: {C5_SHARDED_TOTAL} QPs in total, split by QP index over {world} rank(s)",
            "scaling": "strong", "value": value, "unit": "QP/s", "ms": float(ms.item()), "qps_per_rank": int(mine),
            "status_counts": [int(v) for v in counts.tolist()]}


def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    from mgqp_b200 import workloads
    name = args.workload
    nthreads = host_threads()
    sample = CPU_SAMPLE[name][0]
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

    # The timed region is K launches of ~2 ms: shorter than nvidia-smi's start-up and hardly longer than its sampling
    # period.  So the sampler starts first and the SAME launches run untimed until its first sample is in (pre-roll),
    # and again for ~120 ms after the timed region (post-roll): the clocks reported are those of this load, taken from
    # the start of the timed region on.
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_pre = time.perf_counter()
    while sampler.proc and not sampler.lines and time.perf_counter() - t_pre < 5.0:
        mgqp_b200.solve_quadprog_batched(*dev_in, flags=args.flags, out=out)
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        mgqp_b200.solve_quadprog_batched(*dev_in, flags=args.flags, out=out)
    barrier()
    first_sample = sampler.mark()
    l0 = mgqp_b200.kernel_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        mgqp_b200.solve_quadprog_batched(*dev_in, flags=args.flags, out=out)
        ev[i + 1].record()
    barrier()
    launches = mgqp_b200.kernel_launch_count() - l0
    total_ms = ev[0].elapsed_time(ev[-1])
    t_post = time.perf_counter()
    while time.perf_counter() - t_post < 0.12:   # post-roll (untimed, after the launch count was taken)
        mgqp_b200.solve_quadprog_batched(*dev_in, flags=args.flags, out=out)
        torch.cuda.synchronize()
    clocks = sampler.stop(first_sample)  # every 20 ms from the start of the timed region through the post-roll
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    value = sharding.whole_job_rate(B * args.steps, total_ms / 1e3)  # all ranks' QPs / max-over-ranks device time
    ms_per_step = world * B * 1e3 / value
    status = out.status.cpu().numpy()

    # ---- end-to-end leg: host (pinned) buffers through the plugin call, copies inside the timing
    e2e_steps = max(1, args.e2e_steps or args.steps)
    host_np = [t.numpy() for t in host_pinned]
    res = mgqp_b200.QPResult(torch.empty((B, n), dtype=torch.float64).pin_memory().numpy(),
                             torch.empty((B,), dtype=torch.float64).pin_memory().numpy(),
                             torch.empty((B,), dtype=torch.int32).pin_memory().numpy())
    mgqp_b200.solve_quadprog_batched_host(*host_np, device=local_rank, flags=args.flags, out=res)  # warm
    # (the e2e leg is bound by the host cores and the link, and a 20 ms nvidia-smi poll takes driver locks its
    # copies wait on -- measured: 3.8e7 against 4.9e7 QP/s -- so this leg is sampled at a gentler period)
    sampler2 = ClockSampler(local_rank, period_ms=args.e2e_clock_ms) if args.e2e_clock_ms > 0 else None
    if sampler2:
        sampler2.start()
        sampler2.wait_ready()
    barrier()
    first_sample2 = sampler2.mark() if sampler2 else 0
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        mgqp_b200.solve_quadprog_batched_host(*host_np, device=local_rank, flags=args.flags, out=res)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_clocks = sampler2.stop(first_sample2) if sampler2 else None
    e2e_value = sharding.whole_job_rate(B * e2e_steps, e2e_s)
    # bytes that actually crossed the link in the last call (the library counts its own copies): where a
    # chunk's CI is a set of bounds it travels column-compressed (m*9 bytes, mgqp_batch::CI_val) instead
    # of as the dense n*m*8-byte block the caller handed over
    h2d, d2h = mgqp_b200.last_transfer_bytes()
    host_bytes = sum(a.nbytes for a in host_np)
    e2e_same = bool(np.array_equal(res.status, status) and np.array_equal(res.x, out.x.cpu().numpy()))

    peak, peak_src = peaks()
    # ---- BASELINE.json configs[4] at its stated size, strong-scaled over the ranks (every rank takes part)
    c5s = None
    if not args.no_extras and name == "C2":
        del dev_in
        torch.cuda.empty_cache()
        c5s = c5_sharded_record(mgqp_b200, dev, rank, world, dist)
        dev_in = None
    # ---- the in-library multi-device path (N > 1): rank 0's host threads drive one stream pair per device through
    # solve_quadprog_batched_multi while the other ranks wait; host buffers, chunks from one shared queue
    multi = None
    if world > 1 and not args.no_extras and name == "C2":
        dist.barrier()
        if rank == 0:
            try:
                ndev = min(world, torch.cuda.device_count())
                hb = workloads.make("C5", 1_000_000)
                hp = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy() for a in hb.fields()]
                os.environ["MGQP_HOST_THREADS"] = str(host_threads())
                devs = list(range(ndev))
                r = mgqp_b200.solve_quadprog_batched_multi(*hp, devices=devs)      # warm: arenas, streams
                t0 = time.perf_counter()
                r = mgqp_b200.solve_quadprog_batched_multi(*hp, devices=devs, out=r)
                dt = time.perf_counter() - t0
                up, down = mgqp_b200.last_transfer_bytes()
                multi = {"api": "solve_quadprog_batched_multi (one host thread + two streams per device, chunks from a "
                                "shared atomic queue; no collective)", "devices": ndev, "workload": "C5: 1M QPs in pinned "
                         "host buffers, ~10 % infeasible", "value": hb.batch / dt, "unit": "QP/s",
                         "chunks_per_device": mgqp_b200.last_multi_chunks(), "h2d_bytes": int(up), "d2h_bytes": int(down),
                         "status_counts": np.bincount(r.status & 15, minlength=5).tolist()}
            except Exception as e:
                multi = {"error": repr(e)}
        dist.barrier()
    if rank != 0:
        return
    # ---- parity spot check + CPU baseline on a bounded sample of the same inputs (rank 0, N=1 only)
    line_cpu = None
    parity = None
    contenders = None
    if world == 1 and not args.no_cpu:
        ns = min(B, CPU_SAMPLE[name][0])
        sb = b.slice(0, ns)
        nthreads = host_threads()
        rate, kind = cpu_reference_rate(sb, nthreads)
        line_cpu = {"value": rate, "unit": "QP/s", "cores": nthreads, "kind": kind,
                    "sample": f"first {ns} QPs of the step's batch, {nthreads} host threads, "
                              + ("shipped libquadprog.a (GCC 5.4 -O0) via oracle/_ref" if kind == "reference"
                                 else "oracle/gi_oracle.c -O2")}
        contenders = cpu_contenders(b, name)
        parity = parity_vs_oracle(b, out.x[:ns].cpu().numpy(), out.f[:ns].cpu().numpy(), status, ns)
    kern_ms = float(np.mean(step_ms))  # one launch per step: the step IS the kernel (+ a 4-byte memset)
    achieved = B * bytes_per_qp / (kern_ms / 1e3) / 1e9
    kname = kernel_name(mgqp_b200, n, p, m)
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
                   "kernel": kname, "flags": args.flags,
                   "status_counts": np.bincount(status, minlength=5).tolist()},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "QP/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "matches_device_leg": e2e_same, "host_input_bytes_per_step": int(host_bytes),
                "clocks": e2e_clocks,
                "transport": ("CI as one code byte per +-1 column, packed on the host cores, expanded on the device"
                              if h2d < 0.3 * host_bytes else
                              "CI column-compressed on the host cores (row, value)" if h2d < 0.8 * host_bytes
                              else "dense (every field as the caller holds it)"),
                "api": "solve_quadprog_batched_host (C ABI, pinned host buffers, chunked 2-stream pipeline; CI is "
                       "packed on the host cores into (row, value) columns -- one code byte where the value is +-1 -- when every column is a bound and the "
                       "rank has the host threads to outrun the link"
                       + (" -- disabled by --flags 32)" if args.flags & 32 else ")")},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "bytes_per_qp": bytes_per_qp,
                     "bytes_note": "every input once, x/f/status once (SURVEY.md 8d)", "kernel_ms": kern_ms},
    }
    if line_cpu:
        line["cpu_baseline"] = line_cpu
    if contenders:
        line["cpu_baseline"]["contenders"] = contenders
    if parity:
        line["parity_check"] = parity
    if c5s:
        line["c5_sharded"] = c5s
    if multi:
        line["multi_device_library"] = multi
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr):
        try:
            line["roofline"]["traffic"] = json.load(open(tr)).get(name)
        except Exception:
            pass
    # ---- every other BASELINE config, outside the timed regions above (N = 1 only)
    if world == 1 and not args.no_extras and name == "C2":
        configs = {}
        for cname, cB, csteps in EXTRA_CONFIGS:
            try:
                configs[cname] = extra_config_record(mgqp_b200, cname, cB, csteps, dev, peak)
            except Exception as e:  # a failing sub-record must not take the headline line with it
                configs[cname] = {"error": repr(e)}
        try:
            configs["C4"] = c4_record(mgqp_b200, dev, peak, C4_FULL)
        except Exception as e:
            configs["C4"] = {"error": repr(e)}
        try:
            configs["cascade"] = cascade_record(mgqp_b200, dev)
        except Exception as e:
            configs["cascade"] = {"error": repr(e)}
        try:
            configs["C1_latency"] = c1_latency_record(mgqp_b200)
        except Exception as e:
            configs["C1_latency"] = {"error": repr(e)}
        line["configs"] = configs
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
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer leg (default: --steps)")
    ap.add_argument("--e2e-clock-ms", type=int, default=100, help="nvidia-smi period during the e2e leg (0 = no sampling)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the sub-records (other BASELINE configs, C1 latency, c5_sharded) of the default run")
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
