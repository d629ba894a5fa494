"""TEST INFRASTRUCTURE ONLY -- ctypes loaders for the two CPU checkers.

* ``restatement()``  -> oracle/libgi_oracle.so   (gi_oracle.c, the C restatement; kind "port")
* ``reference()``    -> oracle/_ref/libquadprog_ref.so (the reference's SHIPPED solver binary behind
  ref_shim.cc; kind "reference"); ``None`` when it has not been built.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  Nothing under the product package does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(verbose: bool = False) -> None:
    """Compile the checkers (building the checker is not using it)."""
    r = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout, r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def _ptr(a, ty=_dp):
    return a.ctypes.data_as(ty) if a is not None else None


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Restatement:
    def __init__(self):
        path = os.path.join(HERE, "libgi_oracle.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        self.lib.gi_oracle_solve_batch.restype = C.c_int
        self.lib.gi_oracle_solve_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_longlong] + [_dp] * 8 + [
            _ip, _ip, _ip, _ip, C.c_int, C.c_int]
        self.lib.gi_oracle_solve.restype = C.c_int
        self.lib.gi_oracle_solve.argtypes = [C.c_int, C.c_int, C.c_int] + [_dp] * 8 + [_ip, _ip, _ip, C.c_int]

    def solve_batch(self, n, p, m, G, g0, CE, ce0, CI, ci0, max_iter=0, threads=1):
        B = G.shape[0] if G.ndim == 3 else 1
        G, g0, CE, ce0, CI, ci0 = map(_c, (G, g0, CE, ce0, CI, ci0))
        x = np.zeros((B, n))
        f = np.zeros(B)
        status = np.zeros(B, dtype=np.int32)
        iq = np.zeros(B, dtype=np.int32)
        A = np.zeros((B, p + m), dtype=np.int32)
        iters = np.zeros(B, dtype=np.int32)
        self.lib.gi_oracle_solve_batch(n, p, m, B, _ptr(G), _ptr(g0), _ptr(CE), _ptr(ce0), _ptr(CI), _ptr(ci0),
                                       _ptr(x), _ptr(f), _ptr(status, _ip), _ptr(iq, _ip), _ptr(A, _ip),
                                       _ptr(iters, _ip), int(max_iter), int(threads))
        return dict(x=x, f=f, status=status, iq=iq, A=A, iters=iters)

    def solve_one(self, G, g0, CE, ce0, CI, ci0, max_iter=0):
        """Single QP; returns dict incl. 'G_factor' (G as the solver leaves it)."""
        n = len(g0)
        p = len(ce0)
        m = len(ci0)
        Gc = _c(G).copy().reshape(n, n)
        g0, CE, ce0, CI, ci0 = map(_c, (g0, CE, ce0, CI, ci0))
        x = np.zeros(n)
        f = C.c_double(0)
        iq = C.c_int(0)
        it = C.c_int(0)
        A = np.zeros(p + m + 1, dtype=np.int32)
        st = self.lib.gi_oracle_solve(n, p, m, _ptr(Gc), _ptr(g0), _ptr(CE), _ptr(ce0), _ptr(CI), _ptr(ci0),
                                      _ptr(x), C.byref(f), C.byref(iq), _ptr(A, _ip), C.byref(it), int(max_iter))
        return dict(x=x, f=f.value, status=st, iq=iq.value, A=A[: p + m], iters=it.value, G_factor=Gc)


class Reference:
    """The shipped libquadprog.a (GCC 5.4 -O0) behind oracle/ref_shim.cc."""

    def __init__(self, path):
        self.lib = C.CDLL(path)
        self.lib.ref_solve_quadprog.restype = C.c_int
        self.lib.ref_solve_quadprog.argtypes = [C.c_int, C.c_int, C.c_int] + [_dp] * 9 + [C.c_char_p, C.c_int]
        self.lib.ref_solve_quadprog_batch.restype = C.c_int
        self.lib.ref_solve_quadprog_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_longlong] + [_dp] * 8 + [
            _ip, C.c_int]
        self.lib.ref_dim_mismatch_message.restype = C.c_int
        self.lib.ref_dim_mismatch_message.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_int]

    def solve_one(self, G, g0, CE, ce0, CI, ci0):
        n = len(g0)
        p = len(ce0)
        m = len(ci0)
        G, g0, CE, ce0, CI, ci0 = map(_c, (G, g0, CE, ce0, CI, ci0))
        x = np.zeros(n)
        f = C.c_double(0)
        Gout = np.zeros((n, n))
        msg = C.create_string_buffer(512)
        st = self.lib.ref_solve_quadprog(n, p, m, _ptr(G), _ptr(g0), _ptr(CE), _ptr(ce0), _ptr(CI), _ptr(ci0),
                                         _ptr(x), C.byref(f), _ptr(Gout), msg, 512)
        return dict(x=x, f=f.value, status=st, G_factor=Gout, message=msg.value.decode())

    def solve_batch(self, n, p, m, G, g0, CE, ce0, CI, ci0, threads=1):
        B = G.shape[0]
        G, g0, CE, ce0, CI, ci0 = map(_c, (G, g0, CE, ce0, CI, ci0))
        x = np.zeros((B, n))
        f = np.zeros(B)
        status = np.zeros(B, dtype=np.int32)
        self.lib.ref_solve_quadprog_batch(n, p, m, B, _ptr(G), _ptr(g0), _ptr(CE), _ptr(ce0), _ptr(CI), _ptr(ci0),
                                          _ptr(x), _ptr(f), _ptr(status, _ip), int(threads))
        return dict(x=x, f=f, status=status)

    def dim_mismatch_message(self, n, ce_rows):
        msg = C.create_string_buffer(512)
        st = self.lib.ref_dim_mismatch_message(n, ce_rows, msg, 512)
        return st, msg.value.decode()


_rest = None
_ref = None


def restatement() -> Restatement:
    global _rest
    if _rest is None:
        _rest = Restatement()
    return _rest


def reference() -> Reference | None:
    global _ref
    if _ref is None:
        path = os.path.join(HERE, "_ref", "libquadprog_ref.so")
        if not os.path.exists(path):
            if os.path.exists("/root/reference/lib/QuadProgpp/libquadprog.a"):
                build()
            if not os.path.exists(path):
                return None
        _ref = Reference(path)
    return _ref
