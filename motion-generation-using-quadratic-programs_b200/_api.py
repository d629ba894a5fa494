"""Host-side binding of the C ABI in include/mgqp_b200.h.

The names and argument meaning mirror the reference boundary
(include/QuadProgpp/QuadProg++.hh:69-72, called from src/mgqp.cpp:708/:725):

    f = solve_quadprog(G, g0, CE, ce0, CI, ci0, x)

``G`` n*n, ``CE`` n*p, ``CI`` n*m with one constraint per COLUMN; returns the cost, ``inf`` when
infeasible; raises :class:`LogicError` (the reference throws ``std::logic_error``) for dimension
mismatches and a non-positive-definite ``G`` and :class:`RuntimeError` ("Constraints are linearly
dependent") like the reference's ``std::runtime_error``.  ``G`` is overwritten by its Cholesky
factor (QuadProg++.hh:42-45).

There is NO CPU fallback: everything here calls ``libmgqp_b200.so`` (hand-written sm_100a CUDA);
if the library is missing the import of the solver entry points fails loudly.  torch is used only
for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import NamedTuple, Optional

import numpy as np

__all__ = [
    "LogicError", "QPResult", "STATUS", "lib", "lib_path", "solve_quadprog", "solve_quadprog_batched",
    "solve_quadprog_batched_host", "solve_quadprog_batched_multi", "plan", "algorithmic_bytes",
    "kernel_launch_count", "FLAG_FAST_MATH", "FLAG_NO_BULK_COPY", "FLAG_STATIC_SCHEDULE", "FLAG_GROUP_KERNEL", "FLAG_CTA_KERNEL",
    "FLAG_NO_HOST_PACK", "FLAG_RETRY_WITHOUT_CI", "STATUS_RETRIED", "mgqp_batch", "pack_bound_columns", "pack_unit_columns", "last_transfer_bytes",
    "last_multi_chunks", "mgqp_scenarios", "CascadeResult", "solve_cascade_batched", "solve_cascade_batched_host",
]

_HERE = os.path.dirname(os.path.abspath(__file__))
FLAG_FAST_MATH = 1
FLAG_NO_BULK_COPY = 2
FLAG_STATIC_SCHEDULE = 4
FLAG_GROUP_KERNEL = 8
FLAG_CTA_KERNEL = 16
FLAG_NO_HOST_PACK = 32
FLAG_RETRY_WITHOUT_CI = 64
STATUS_RETRIED = 16
STATUS = {0: "ok", 1: "infeasible", 2: "not_pd", 3: "eq_dependent", 4: "iter_cap"}


class LogicError(ValueError):
    """What the reference reports as std::logic_error."""


class mgqp_batch(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("p", C.c_int32), ("m", C.c_int32), ("batch", C.c_int64),
        ("G", C.c_void_p), ("g0", C.c_void_p), ("CE", C.c_void_p), ("ce0", C.c_void_p),
        ("CI", C.c_void_p), ("ci0", C.c_void_p),
        ("sG", C.c_int64), ("sg0", C.c_int64), ("sCE", C.c_int64), ("sce0", C.c_int64),
        ("sCI", C.c_int64), ("sci0", C.c_int64),
        ("x", C.c_void_p), ("f", C.c_void_p), ("status", C.c_void_p), ("iq", C.c_void_p),
        ("A", C.c_void_p), ("iters", C.c_void_p),
        ("max_iter", C.c_int32), ("flags", C.c_int32),
        ("CI_val", C.c_void_p), ("CI_row", C.c_void_p), ("sCIv", C.c_int64), ("sCIr", C.c_int64),
    ]


class mgqp_scenarios(C.Structure):
    _fields_ = ([("batch", C.c_int64)]
                + [(k, C.c_void_p) for k in ("q", "qd", "M", "h", "J", "Jd", "pos_des", "pos_cur", "vel_des", "vel_cur",
                                             "acc_des", "acc_cur", "q1_des")]
                + [(k, C.c_float) for k in ("kp_task", "kd_task", "kp_joint", "kd_joint")]
                + [(k, C.c_float * 7) for k in ("acc_p", "acc_n", "tau_p", "tau_n", "q_sup", "q_inf")]
                + [("joint_task", C.c_int32)]
                + [(k, C.c_void_p) for k in ("tracking", "torques", "status0", "status2", "tap_CE0", "tap_ce0",
                                             "tap_ci0", "tap_Z")])


class CascadeResult(NamedTuple):
    tracking: object       # [B,14] float32: [qdd; tau] as solveNextHierarchy returns it
    torques: object        # [B,7]  float32: tau + h (src/mgqp.cpp:1152)
    status0: object
    status2: object
    CE0: object = None     # taps (device path, taps=True): the level-0 QP as assembled, the projector
    ce0: object = None
    ci0: object = None
    Z: object = None


class QPResult(NamedTuple):
    x: object
    f: object
    status: object
    iq: object = None
    A: object = None
    iters: object = None


_lib = None


def lib_path() -> str:
    # MGQP_B200_LIB: developer override used by tools/ab_build.sh to time kernel variants side by side
    return os.environ.get("MGQP_B200_LIB") or os.path.join(_HERE, "libmgqp_b200.so")


def lib():
    """The CUDA library.  Raises if it has not been built (``python __graft_entry__.py``)."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "mgqp_b200 has no CPU fallback.")
        L = C.CDLL(path)
        L.solve_quadprog_batched.restype = C.c_int
        L.solve_quadprog_batched.argtypes = [C.POINTER(mgqp_batch), C.c_int, C.c_void_p]
        L.solve_quadprog_batched_host.restype = C.c_int
        L.solve_quadprog_batched_host.argtypes = [C.POINTER(mgqp_batch), C.c_int, C.c_int64]
        L.solve_quadprog_batched_multi.restype = C.c_int
        L.solve_quadprog_batched_multi.argtypes = [C.POINTER(mgqp_batch), C.POINTER(C.c_int), C.c_int, C.c_int64]
        L.mgqp_solve_one.restype = C.c_int
        L.mgqp_solve_one.argtypes = [C.c_int32] * 3 + [C.c_void_p] * 6 + [C.c_void_p, C.POINTER(C.c_double),
                                                                           C.POINTER(C.c_int32), C.c_void_p, C.c_int]
        L.mgqp_status_string.restype = C.c_char_p
        L.mgqp_status_string.argtypes = [C.c_int]
        L.mgqp_last_error.restype = C.c_char_p
        L.mgqp_device_count.restype = C.c_int
        L.mgqp_kernel_launch_count.restype = C.c_int64
        L.mgqp_plan.restype = C.c_int
        L.mgqp_plan.argtypes = [C.c_int32] * 3 + [C.POINTER(C.c_int32)] * 2
        L.mgqp_algorithmic_bytes.restype = C.c_int64
        L.mgqp_algorithmic_bytes.argtypes = [C.c_int32] * 3
        L.mgqp_pack_bound_columns.restype = C.c_int
        L.mgqp_pack_bound_columns.argtypes = [C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                              C.c_void_p, C.c_int64, C.c_int]
        L.mgqp_pack_unit_columns.restype = C.c_int
        L.mgqp_pack_unit_columns.argtypes = [C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                             C.c_int]
        L.mgqp_cascade_batched.restype = C.c_int
        L.mgqp_cascade_batched.argtypes = [C.POINTER(mgqp_scenarios), C.c_int, C.c_void_p]
        L.mgqp_cascade_batched_host.restype = C.c_int
        L.mgqp_cascade_batched_host.argtypes = [C.POINTER(mgqp_scenarios), C.c_int, C.c_int64]
        L.mgqp_last_transfer_bytes.restype = None
        L.mgqp_last_transfer_bytes.argtypes = [C.POINTER(C.c_int64)] * 2
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        msg = lib().mgqp_last_error().decode()
        if rc == -1:
            raise ValueError(f"mgqp_b200: invalid argument: {msg}")
        if rc == -2:
            raise NotImplementedError(f"mgqp_b200: {msg}")
        raise RuntimeError(f"mgqp_b200: error {rc}: {msg}")


def plan(n: int, p: int, m: int):
    """(lanes per QP, warps per CTA, dynamic smem bytes per CTA) the launcher would use."""
    smem, warps = C.c_int32(0), C.c_int32(0)
    gw = lib().mgqp_plan(n, p, m, C.byref(smem), C.byref(warps))
    return gw, warps.value, smem.value


def algorithmic_bytes(n: int, p: int, m: int) -> int:
    return int(lib().mgqp_algorithmic_bytes(n, p, m))


def kernel_launch_count() -> int:
    return int(lib().mgqp_kernel_launch_count())


def last_transfer_bytes():
    """(host->device, device->host) bytes moved by this thread's last host-buffer call."""
    up, down = C.c_int64(0), C.c_int64(0)
    lib().mgqp_last_transfer_bytes(C.byref(up), C.byref(down))
    return up.value, down.value


def last_multi_chunks():
    """Chunks each device took in this thread's last solve_quadprog_batched_multi call."""
    buf = (C.c_int64 * 16)()
    lib().mgqp_last_multi_chunks.restype = C.c_int
    lib().mgqp_last_multi_chunks.argtypes = [C.POINTER(C.c_int64), C.c_int]
    n = lib().mgqp_last_multi_chunks(buf, 16)
    return [int(buf[i]) for i in range(n)]


def pack_bound_columns(CI, threads: int = 0):
    """Column-compress ``CI [B,n,m]`` (host array): returns ``(val [B,m] float64, row [B,m] uint8)`` or
    ``None`` if some column has more than one nonzero coefficient (mgqp_batch::CI_val in the header)."""
    CI = np.ascontiguousarray(CI, dtype=np.float64)
    B, n, m = CI.shape
    val = np.empty((B, m))
    row = np.empty((B, m), dtype=np.uint8)
    ok = lib().mgqp_pack_bound_columns(n, m, B, CI.ctypes.data, n * m, val.ctypes.data, m, row.ctypes.data, m, threads)
    return (val, row) if ok else None


def pack_unit_columns(CI, threads: int = 0):
    """One code byte per column of ``CI [B,n,m]`` when every nonzero coefficient is exactly +-1.0 and no column has
    two (bits 0-5 the row, bit 7 set for -1.0, 0x40 for an all-zero column): ``code [B,m] uint8``, or ``None``."""
    CI = np.ascontiguousarray(CI, dtype=np.float64)
    B, n, m = CI.shape
    code = np.empty((B, m), dtype=np.uint8)
    ok = lib().mgqp_pack_unit_columns(n, m, B, CI.ctypes.data, n * m, code.ctypes.data, m, threads)
    return code if ok else None


# ------------------------------------------------------------------ device-resident batches
def solve_quadprog_batched(G, g0, CE, ce0, CI, ci0, *, want_active_set: bool = False, max_iter: int = 0,
                           flags: int = 0, stream=None, out: Optional[QPResult] = None,
                           CI_val=None, CI_row=None) -> QPResult:
    """Solve a batch of QPs held in CUDA tensors (float64).

    Shapes: ``G [B,n,n]``, ``g0 [B,n]``, ``CE [B,n,p]``, ``ce0 [B,p]``, ``CI [B,n,m]``, ``ci0 [B,m]``;
    a leading dimension of 1 broadcasts that field to the whole batch.  Asynchronous on ``stream``
    (default: torch's current stream).  Returns torch tensors on the same device.

    ``CI_val [B,m]`` (float64) / ``CI_row [B,m]`` (uint8), if given, replace ``CI`` by its
    column-compressed form (every inequality a bound on one variable; ``CI`` may then be ``None``).
    """
    import torch

    if not G.is_cuda:
        raise ValueError("solve_quadprog_batched takes CUDA tensors; use solve_quadprog_batched_host for host arrays")
    dev = G.device
    n = G.shape[-1]
    p = ce0.shape[-1]
    m = ci0.shape[-1]
    if CI_val is not None:
        if CI_row is None or CI_val.dtype != torch.float64 or CI_row.dtype != torch.uint8:
            raise ValueError("CI_val (float64) and CI_row (uint8) go together")
        if tuple(CI_val.shape[1:]) != (m,) or tuple(CI_row.shape[1:]) != (m,):
            raise LogicError(f"The compressed CI is incompatible (expecting [B, {m}])")
        if CI is None:
            CI = torch.empty((1, n, m), dtype=torch.float64, device=dev)  # never read
    B = max(t.shape[0] for t in (G, g0, CE, ce0, CI, ci0))
    if CI_val is not None and (CI_val.shape[0] not in (1, B) or CI_row.shape[0] not in (1, B)):
        raise ValueError("leading dimensions of CI_val / CI_row must be B or 1")
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    sptr = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
    # Temporaries (contiguous copies) are produced on torch's current stream and handed back to its caching
    # allocator when this function returns; on a foreign stream nothing would order either against the
    # asynchronous solver launch.  There the caller must pass contiguous tensors.
    foreign = sptr != torch.cuda.current_stream(dev).cuda_stream

    def _contig(t):
        if t.numel() == 0 or t[0].is_contiguous():
            return t
        if foreign:
            raise ValueError("non-contiguous input with a stream that is not torch's current stream: "
                             "make the tensor contiguous on that stream first")
        return t.contiguous()

    if CI_val is not None:
        CI_val, CI_row = _contig(CI_val), _contig(CI_row)
    _validate_shapes(n, p, m, tuple(G.shape[1:]), tuple(g0.shape[1:]), tuple(CE.shape[1:]), tuple(ce0.shape[1:]),
                     tuple(CI.shape[1:]), tuple(ci0.shape[1:]))
    ts = []
    for t in (G, g0, CE, ce0, CI, ci0):
        if t.dtype != torch.float64:
            raise ValueError("all inputs must be float64")
        if t.shape[0] not in (1, B):
            raise ValueError("leading dimensions must be B or 1")
        ts.append(_contig(t))
    G, g0, CE, ce0, CI, ci0 = ts
    if out is None:
        x = torch.empty((B, n), dtype=torch.float64, device=dev)
        f = torch.empty((B,), dtype=torch.float64, device=dev)
        status = torch.empty((B,), dtype=torch.int32, device=dev)
        iq = A = iters = None
        if want_active_set:
            iq = torch.empty((B,), dtype=torch.int32, device=dev)
            A = torch.empty((B, p + m), dtype=torch.int32, device=dev)
            iters = torch.empty((B,), dtype=torch.int32, device=dev)
        out = QPResult(x, f, status, iq, A, iters)
    if B == 0:
        return out
    if p > n:
        raise ValueError(f"p = {p} equalities on n = {n} variables (the reference indexes R out of bounds there)")
    d = mgqp_batch()
    d.n, d.p, d.m, d.batch = n, p, m, B
    nat = (n * n, n, n * p, p, n * m, m)
    for name, t, na in zip(("G", "g0", "CE", "ce0", "CI", "ci0"), ts, nat):
        setattr(d, name, t.data_ptr() if t.numel() else out.x.data_ptr())
        setattr(d, "s" + name, (0 if t.shape[0] == 1 and B > 1 else (t.stride(0) if na else 0)))
    d.x, d.f, d.status = out.x.data_ptr(), out.f.data_ptr(), out.status.data_ptr()
    d.iq = out.iq.data_ptr() if out.iq is not None else None
    d.A = out.A.data_ptr() if out.A is not None else None
    d.iters = out.iters.data_ptr() if out.iters is not None else None
    d.max_iter, d.flags = max_iter, flags
    if CI_val is not None and m:
        d.CI_val, d.CI_row = CI_val.data_ptr(), CI_row.data_ptr()
        d.sCIv = 0 if CI_val.shape[0] == 1 and B > 1 else m
        d.sCIr = 0 if CI_row.shape[0] == 1 and B > 1 else m
    _check(lib().solve_quadprog_batched(C.byref(d), dev.index if dev.index is not None else 0, C.c_void_p(sptr)))
    return out


# ------------------------------------------------------------------ host buffers
def _host_desc(G, g0, CE, ce0, CI, ci0, want_active_set, max_iter, flags, outs=None):
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (G, g0, CE, ce0, CI, ci0)]
    G, g0, CE, ce0, CI, ci0 = arrs
    names = ("G", "g0", "CE", "ce0", "CI", "ci0")
    for name, a, nd in zip(names, arrs, (3, 2, 3, 2, 3, 2)):
        if a.ndim != nd:
            raise ValueError(f"{name} must have {nd} dimensions ([B, ...]; a leading 1 broadcasts), got shape {a.shape}")
    n = g0.shape[1]
    p = ce0.shape[1]
    m = ci0.shape[1]
    _validate_shapes(n, p, m, G.shape[1:], g0.shape[1:], CE.shape[1:], ce0.shape[1:], CI.shape[1:], ci0.shape[1:])
    # The C side walks every field with batch * stride: a field shorter than the batch would be read
    # past its end.  Leading dimension B, or 1 = one instance for the whole batch (stride 0).
    B = max(a.shape[0] for a in arrs)
    for name, a in zip(names, arrs):
        if a.shape[0] not in (1, B):
            raise ValueError(f"{name} has leading dimension {a.shape[0]}; expecting the batch size {B} or 1 (broadcast)")
    if outs is None:
        x = np.empty((B, n))
        f = np.empty(B)
        status = np.empty(B, dtype=np.int32)
        iq = A = iters = None
        if want_active_set:
            iq = np.empty(B, dtype=np.int32)
            A = np.empty((B, p + m), dtype=np.int32)
            iters = np.empty(B, dtype=np.int32)
        outs = QPResult(x, f, status, iq, A, iters)
    else:
        if outs.x.shape != (B, n) or outs.f.shape != (B,) or outs.status.shape != (B,):
            raise ValueError(f"out buffers do not match the batch ([{B}, {n}], [{B}], [{B}])")
    d = mgqp_batch()
    d.n, d.p, d.m, d.batch = n, p, m, B
    nat = (n * n, n, n * p, p, n * m, m)
    dummy = outs.x
    for name, a, na in zip(names, arrs, nat):
        setattr(d, name, a.ctypes.data if a.size else dummy.ctypes.data)
        setattr(d, "s" + name, 0 if (a.shape[0] == 1 and B > 1) else na)
    d.x, d.f, d.status = outs.x.ctypes.data, outs.f.ctypes.data, outs.status.ctypes.data
    d.iq = outs.iq.ctypes.data if outs.iq is not None else None
    d.A = outs.A.ctypes.data if outs.A is not None else None
    d.iters = outs.iters.ctypes.data if outs.iters is not None else None
    d.max_iter, d.flags = max_iter, flags
    return d, outs, arrs


def solve_quadprog_batched_host(G, g0, CE, ce0, CI, ci0, *, want_active_set=False, max_iter=0, flags=0,
                                device: int = 0, chunk: int = 0, out: Optional[QPResult] = None) -> QPResult:
    """Host (numpy) arrays in, numpy arrays out; H2D / solve / D2H pipelined inside the library."""
    d, outs, keep = _host_desc(G, g0, CE, ce0, CI, ci0, want_active_set, max_iter, flags, out)
    _check(lib().solve_quadprog_batched_host(C.byref(d), device, chunk))
    return outs


def solve_quadprog_batched_multi(G, g0, CE, ce0, CI, ci0, *, devices, want_active_set=False, max_iter=0, flags=0,
                                 chunk: int = 0, out: Optional[QPResult] = None) -> QPResult:
    """Host arrays sharded by QP index over several GPUs, one host thread per device (no collective)."""
    d, outs, keep = _host_desc(G, g0, CE, ce0, CI, ci0, want_active_set, max_iter, flags, out)
    devs = (C.c_int * len(devices))(*devices)
    _check(lib().solve_quadprog_batched_multi(C.byref(d), devs, len(devices), chunk))
    return outs


# ------------------------------------------------------------------ the solver's callers, batched
_SCEN_SHAPES = {"q": (7,), "qd": (7,), "M": (7, 7), "h": (7,), "J": (3, 7), "Jd": (3, 7), "pos_des": (3,), "pos_cur": (3,),
                "vel_des": (3,), "vel_cur": (3,), "acc_des": (3,), "acc_cur": (3,), "q1_des": ()}


def _scenario_desc(state, config, B):
    from .workloads import SHIPPED_CONFIG
    cfg = SHIPPED_CONFIG if config is None else config
    d = mgqp_scenarios()
    d.batch = B
    for k in ("kp_task", "kd_task", "kp_joint", "kd_joint"):
        setattr(d, k, float(cfg[k]))
    for k in ("acc_p", "acc_n", "tau_p", "tau_n", "q_sup", "q_inf"):
        v = np.asarray(cfg[k], dtype=np.float32).reshape(-1)
        if v.size != 7:
            raise ValueError(f"config[{k!r}] must have 7 entries")
        setattr(d, k, (C.c_float * 7)(*[float(x) for x in v]))
    d.joint_task = int(cfg["joint_task"])
    return d


def solve_cascade_batched(state, config=None, *, taps: bool = False, stream=None) -> CascadeResult:
    """QP assembly (updateHook, src/mgqp.cpp:886-1143) + priority cascade (solveNextHierarchy, :751-869) for a batch
    of robot states held in CUDA float32 tensors: ``state`` maps the field names of ``mgqp_scenarios`` (q, qd, M, h, J,
    Jd, pos_des, pos_cur, vel_des, vel_cur, acc_des, acc_cur, q1_des) to ``[B, ...]`` tensors; ``config`` holds the
    gains and limits (default: as shipped).  Asynchronous on ``stream`` (default: torch's current stream)."""
    import torch

    q = state["q"]
    if not q.is_cuda:
        raise ValueError("solve_cascade_batched takes CUDA tensors; use solve_cascade_batched_host for host arrays")
    dev, B = q.device, q.shape[0]
    d = _scenario_desc(state, config, B)
    keep = []
    for k, shp in _SCEN_SHAPES.items():
        t = state[k]
        if t.dtype != torch.float32 or tuple(t.shape) != (B,) + shp or not t.is_contiguous() or t.device != dev:
            raise ValueError(f"state[{k!r}] must be a contiguous float32 CUDA tensor of shape {(B,) + shp}")
        setattr(d, k, t.data_ptr())
        keep.append(t)
    f32 = dict(dtype=torch.float32, device=dev)
    out = CascadeResult(torch.empty((B, 14), **f32), torch.empty((B, 7), **f32),
                        torch.empty((B,), dtype=torch.int32, device=dev), torch.empty((B,), dtype=torch.int32, device=dev),
                        *((torch.empty((B, 14, 10), dtype=torch.float64, device=dev),
                           torch.empty((B, 10), dtype=torch.float64, device=dev),
                           torch.empty((B, 28), dtype=torch.float64, device=dev),
                           torch.empty((B, 14, 14), **f32)) if taps else ()))
    d.tracking, d.torques = out.tracking.data_ptr(), out.torques.data_ptr()
    d.status0, d.status2 = out.status0.data_ptr(), out.status2.data_ptr()
    if taps:
        d.tap_CE0, d.tap_ce0, d.tap_ci0, d.tap_Z = (out.CE0.data_ptr(), out.ce0.data_ptr(), out.ci0.data_ptr(),
                                                    out.Z.data_ptr())
    if B == 0:
        return out
    if stream is None:
        stream = torch.cuda.current_stream(dev)
    sptr = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
    _check(lib().mgqp_cascade_batched(C.byref(d), dev.index if dev.index is not None else 0, C.c_void_p(sptr)))
    return out


def solve_cascade_batched_host(state, config=None, *, device: int = 0, chunk: int = 0,
                               out: Optional[CascadeResult] = None) -> CascadeResult:
    """Host (numpy float32) arrays in, numpy arrays out: ~0.5 KB of robot state up and 84 bytes of command down per
    scenario, the QPs themselves are assembled, solved and projected on the device."""
    arrs = {k: np.ascontiguousarray(state[k], dtype=np.float32) for k in _SCEN_SHAPES}
    B = arrs["q"].shape[0]
    d = _scenario_desc(state, config, B)
    for k, shp in _SCEN_SHAPES.items():
        if arrs[k].shape != (B,) + shp:
            raise ValueError(f"state[{k!r}] must have shape {(B,) + shp}, got {arrs[k].shape}")
        setattr(d, k, arrs[k].ctypes.data)
    if out is None:
        out = CascadeResult(np.empty((B, 14), np.float32), np.empty((B, 7), np.float32), np.empty(B, np.int32),
                            np.empty(B, np.int32))
    d.tracking, d.torques = out.tracking.ctypes.data, out.torques.ctypes.data
    d.status0, d.status2 = out.status0.ctypes.data, out.status2.ctypes.data
    _check(lib().mgqp_cascade_batched_host(C.byref(d), device, chunk))
    return out


# ------------------------------------------------------------------ the reference's own call
def _validate_shapes(n, p, m, sG, sg0, sCE, sce0, sCI, sci0):
    """The five dimension checks of solve_quadprog, with the reference's exception texts
    (.rodata of QuadProg++.o)."""
    sG, sCE, sCI = tuple(sG), tuple(sCE), tuple(sCI)
    if sG[0] != sG[1]:
        raise LogicError(f"The matrix G is not a squared matrix ({sG[0]} x {sG[1]})")
    if sCE[0] != n:
        raise LogicError(f"The matrix CE is incompatible (incorrect number of rows {sCE[0]} , expecting {n})")
    if tuple(sce0)[0] != sCE[1]:
        raise LogicError(f"The vector ce0 is incompatible (incorrect dimension {tuple(sce0)[0]}, expecting {sCE[1]})")
    if sCI[0] != n:
        raise LogicError(f"The matrix CI is incompatible (incorrect number of rows {sCI[0]} , expecting {n})")
    if tuple(sci0)[0] != sCI[1]:
        raise LogicError(f"The vector ci0 is incompatible (incorrect dimension {tuple(sci0)[0]}, expecting {sCI[1]})")


def _fmt_double(v: float) -> str:
    """std::ostream << double with default precision (6 significant digits, %g)."""
    return "%g" % v


def solve_quadprog(G, g0, CE, ce0, CI, ci0, x=None, *, device: int = 0) -> float:
    """Single QP, reference semantics (QuadProg++.hh:69-72).

    ``G`` (float64 ndarray, n*n) is overwritten by its Cholesky factor; ``x`` (optional float64
    ndarray of length n) receives the solution (also kept as ``solve_quadprog.last_x``).
    Returns the cost, ``math.inf`` if infeasible.
    """
    G_arr = np.asarray(G)
    g0 = np.ascontiguousarray(g0, dtype=np.float64)
    n = G_arr.shape[1] if G_arr.ndim == 2 else 0
    CE = np.ascontiguousarray(CE, dtype=np.float64)
    CI = np.ascontiguousarray(CI, dtype=np.float64)
    if G_arr.ndim != 2 or CE.ndim != 2 or CI.ndim != 2:
        raise LogicError("G, CE and CI must be 2-D (n*n, n*p, n*m; use shape (n, 0) for an empty block)")
    ce0 = np.ascontiguousarray(ce0, dtype=np.float64).reshape(-1)
    ci0 = np.ascontiguousarray(ci0, dtype=np.float64).reshape(-1)
    _validate_shapes(n, CE.shape[1], CI.shape[1], G_arr.shape, g0.shape, CE.shape, ce0.shape, CI.shape, ci0.shape)
    Gc = np.ascontiguousarray(G_arr, dtype=np.float64)
    p, m = CE.shape[1], CI.shape[1]
    xs = np.zeros(n)
    f = C.c_double(0.0)
    st = C.c_int32(0)
    Gf = np.empty((n, n))
    _check(lib().mgqp_solve_one(n, p, m, Gc.ctypes.data, g0.ctypes.data, CE.ctypes.data if CE.size else None,
                                ce0.ctypes.data if ce0.size else None, CI.ctypes.data if CI.size else None,
                                ci0.ctypes.data if ci0.size else None, xs.ctypes.data, C.byref(f), C.byref(st),
                                Gf.ctypes.data, device))
    if st.value == 2:
        raise LogicError("Error in cholesky decomposition, sum: " + _fmt_double(f.value))
    if isinstance(G, np.ndarray) and G.dtype == np.float64:
        G[...] = Gf  # "The matrix G is modified within the function" (QuadProg++.hh:42-45)
    if x is not None:
        x[...] = xs
    solve_quadprog.last_x = xs
    solve_quadprog.last_status = st.value
    if st.value == 3:
        raise RuntimeError("Constraints are linearly dependent")
    if st.value == 1:
        return math.inf
    return f.value
