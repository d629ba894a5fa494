"""Shared comparison helpers for the parity tests."""
import ctypes as C

import numpy as np

RTOL = 1e-9  # BASELINE.json north_star: "x and the objective within 1e-9 relative in fp64"


def emu_solve(lib, b, gw, fma=0, max_iter=0, threads=4):
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    B, n, p, m = b.batch, b.n, b.p, b.m
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in b.fields()]
    x = np.zeros((B, n)); f = np.zeros(B)
    st = np.zeros(B, np.int32); iq = np.zeros(B, np.int32); A = np.zeros((B, p + m), np.int32); it = np.zeros(B, np.int32)
    rc = lib.emu_solve_batch(n, p, m, B, *[a.ctypes.data_as(dp) for a in arrs], x.ctypes.data_as(dp),
                             f.ctypes.data_as(dp), *[a.ctypes.data_as(ip) for a in (st, iq, A, it)], max_iter, gw, fma,
                             threads)
    assert rc == 0
    return dict(x=x, f=f, status=st, iq=iq, A=A, iters=it)


def assert_bit_identical(got, ref, what=""):
    for k in ("status", "iq", "A", "iters"):
        if k in got and got[k] is not None and k in ref:
            assert np.array_equal(got[k], ref[k]), f"{what}: {k} differs at {np.nonzero(np.atleast_1d(got[k] != ref[k]))[0][:5]}"
    for k in ("x", "f"):
        _assert_same_bits(got[k], ref[k], f"{what}: {k}")


def _assert_same_bits(a, b, what):
    """Equality of the fp64 BIT patterns (so -0.0 != +0.0), any NaN matching any NaN."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    same = (a.view(np.int64) == b.view(np.int64)) | (np.isnan(a) & np.isnan(b))
    assert same.all(), f"{what} not bit-identical at {np.argwhere(~same)[:5].tolist()}"


def assert_parity(got, ref, what="", rtol=RTOL):
    """The north-star gate: identical verdicts and active sets; x, f within 1e-9 relative."""
    assert np.array_equal(got["status"], ref["status"]), f"{what}: verdicts differ"
    if got.get("iq") is not None and "iq" in ref:
        assert np.array_equal(got["iq"], ref["iq"]), f"{what}: active-set sizes differ"
        assert np.array_equal(got["A"], ref["A"]), f"{what}: active sets differ"
    ok = ref["status"] == 0
    xs = np.maximum(1.0, np.abs(ref["x"][ok]).max(axis=1, keepdims=True)) if ok.any() else 1.0
    assert np.all(np.abs(got["x"][ok] - ref["x"][ok]) <= rtol * xs), f"{what}: x outside {rtol} relative"
    fs = np.maximum(1.0, np.abs(ref["f"][ok]))
    assert np.all(np.abs(got["f"][ok] - ref["f"][ok]) <= rtol * fs), f"{what}: f outside {rtol} relative"
    inf = ref["status"] == 1
    assert np.all(np.isposinf(got["f"][inf])), f"{what}: infeasible QPs must return +inf"
