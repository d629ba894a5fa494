"""CPU: the CTA-per-QP DEVICE solver body (csrc/gi_large.cuh) compiled for the host and run as a
thread block of fibers (tests/emu), against the oracle: bit-identical x, f, verdict, active set and
step count.  Covers the three-pass Givens sweep and the block-wide argmin / broadcast control flow."""
import ctypes as C

import numpy as np
import pytest

from conftest import default_max_iter
from helpers import assert_bit_identical
from mgqp_b200 import workloads
from test_emu_lockstep import _degenerate_batch


def large_solve(lib, b, max_iter=0, nthreads=64):
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.emu_large_solve_batch.restype = C.c_int
    lib.emu_large_solve_batch.argtypes = [C.c_int] * 3 + [C.c_longlong] + [dp] * 8 + [ip] * 4 + [C.c_int] * 4
    B, n, p, m = b.batch, b.n, b.p, b.m
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in b.fields()]
    x = np.zeros((B, n)); f = np.zeros(B)
    st = np.zeros(B, np.int32); iq = np.zeros(B, np.int32); A = np.zeros((B, p + m), np.int32); it = np.zeros(B, np.int32)
    rc = lib.emu_large_solve_batch(n, p, m, B, *[a.ctypes.data_as(dp) for a in arrs], x.ctypes.data_as(dp),
                                   f.ctypes.data_as(dp), *[a.ctypes.data_as(ip) for a in (st, iq, A, it)], max_iter,
                                   nthreads, 0, 4)
    assert rc == 0
    return dict(x=x, f=f, status=st, iq=iq, A=A, iters=it)


def _oracle(restatement, b, max_iter=0):
    return restatement.solve_batch(b.n, b.p, b.m, *b.fields(), max_iter=max_iter or default_max_iter(b.p, b.m), threads=4)


@pytest.mark.parametrize("name,B", [("C2", 64), ("C3", 48), ("S0", 48)])
def test_benchmark_shapes(emu_lib, restatement, name, B):
    b = workloads.make(name, B)
    assert_bit_identical(large_solve(emu_lib, b), _oracle(restatement, b), name)


@pytest.mark.parametrize("n,p,m,nt", [(40, 3, 90, 64), (33, 0, 70, 64), (50, 5, 160, 32), (70, 0, 150, 128), (3, 1, 0, 64)])
def test_shapes_beyond_the_warp_kernels(emu_lib, restatement, n, p, m, nt):
    b = workloads.make_random(12, n, p, m, seed=n * 100 + m, box=(m >= 2 * n))
    assert_bit_identical(large_solve(emu_lib, b, nthreads=nt), _oracle(restatement, b), f"R{n}_{p}_{m}")


def test_degenerate_and_failures(emu_lib, restatement):
    b = _degenerate_batch(48)
    for cap in (7, 60):
        assert_bit_identical(large_solve(emu_lib, b, cap), _oracle(restatement, b, cap), f"cap{cap}")
    b = workloads.make_random(64, 6, 2, 14, seed=11, box=True)
    b.CE[0::4, :, 1] = -0.5 * b.CE[0::4, :, 0]
    b.G[1::8] -= 4.0 * np.eye(6)
    b.ci0[2::8, 0] = -3.0; b.ci0[2::8, 6] = 1.0
    ref = _oracle(restatement, b)
    got = large_solve(emu_lib, b)
    assert np.array_equal(got["status"], ref["status"])
    keep = ref["status"] != 2
    for k in ("x", "iq", "iters", "A"):
        assert np.array_equal(got[k][keep], ref[k][keep]), k
    assert np.array_equal(got["f"], ref["f"], equal_nan=True)


def test_poisoned_factor_takes_the_full_loops(emu_lib, restatement):
    """inf / NaN inside G: the upper-triangular shortcut of the J computation must not be used, the
    poison has to spread exactly as in the reference."""
    b = workloads.make_random(12, 40, 2, 60, seed=77)
    b.G[1, 3, 4] = b.G[1, 4, 3] = np.inf
    b.G[2, 5, 5] = np.nan
    b.G[3, 39, 39] = np.inf
    ref = _oracle(restatement, b)
    got = large_solve(emu_lib, b)
    assert np.array_equal(got["status"], ref["status"])
    keep = ref["status"] != 2
    assert np.array_equal(got["x"][keep], ref["x"][keep], equal_nan=True)
    assert np.array_equal(got["f"], ref["f"], equal_nan=True)
    assert np.array_equal(got["iters"][keep], ref["iters"][keep])


def test_horizon_stacked_structure_uses_the_sparse_scan(emu_lib, restatement):
    """The C4 construction (identity and integrator blocks: a few nonzeros per column) at a size the emulator can
    run: the ELL form of CI must give bit-identical scans, also with -0.0 / NaN coefficients and a poisoned iterate."""
    b = workloads.make_c4(6, seed=5, dof=3, steps=14)          # n = 42, m = 252
    b.CI[1, 4, 7] = -0.0
    b.CI[2, 5, 9] = np.nan
    b.g0[3, 2] = np.inf                                        # x0 is not finite: the scan falls back to the dense sums
    assert_bit_identical(large_solve(emu_lib, b), _oracle(restatement, b), "mpc")
    b = workloads.make_c4(4, seed=6, dof=7, steps=12)          # n = 84 > 64: columns with more than 64 nonzeros stay dense
    b.CI[:, :, 5] = np.random.default_rng(1).standard_normal((4, 84))
    assert_bit_identical(large_solve(emu_lib, b, nthreads=128), _oracle(restatement, b), "mpc84")
