"""CPU: the one-QP-per-lane DEVICE solver body (csrc/gi_lane.cuh, incl. the bound-column
classification it does while staging CI) compiled for the host, against the oracle: bit-identical
x, f, verdict, active set and step count."""
import ctypes as C

import numpy as np
import pytest

from conftest import default_max_iter
from helpers import assert_bit_identical
from mgqp_b200 import workloads
from test_emu_lockstep import _degenerate_batch


def lane_solve(lib, b, max_iter=0, fma=0, compressed=None):
    """compressed = (val [B,m], row [B,m] uint8): solve from the column-compressed form of CI (the dense
    block is then withheld from the solver altogether)."""
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    lib.emu_lane_solve_batch_c.restype = C.c_int
    lib.emu_lane_solve_batch_c.argtypes = [C.c_int] * 3 + [C.c_longlong] + [dp] * 8 + [ip] * 4 + [C.c_int] * 2 + [C.c_void_p] * 2
    B, n, p, m = b.batch, b.n, b.p, b.m
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in b.fields()]
    x = np.zeros((B, n)); f = np.zeros(B)
    st = np.zeros(B, np.int32); iq = np.zeros(B, np.int32); A = np.zeros((B, p + m), np.int32); it = np.zeros(B, np.int32)
    ptrs = [a.ctypes.data_as(dp) for a in arrs]
    civ = cir = None
    if compressed is not None:
        val = np.ascontiguousarray(compressed[0], dtype=np.float64)
        row = np.ascontiguousarray(compressed[1], dtype=np.uint8)
        civ, cir = val.ctypes.data, row.ctypes.data
        ptrs[4] = None
    rc = lib.emu_lane_solve_batch_c(n, p, m, B, *ptrs, x.ctypes.data_as(dp),
                                    f.ctypes.data_as(dp), *[a.ctypes.data_as(ip) for a in (st, iq, A, it)], max_iter, fma,
                                    civ, cir)
    assert rc == 0
    return dict(x=x, f=f, status=st, iq=iq, A=A, iters=it)


def _compress(CI):
    """numpy restatement of the column compression (test-side, independent of the library's packer)."""
    nz = (CI != 0) | np.isnan(CI)
    assert (nz.sum(axis=1) <= 1).all(), "not a bound-only batch"
    row = nz.argmax(axis=1).astype(np.uint8)                      # [B, m]
    val = np.take_along_axis(CI, row[:, None, :].astype(np.int64), axis=1)[:, 0, :]
    return np.where(nz.any(axis=1), val, 0.0), row


def _oracle(restatement, b, max_iter=0):
    return restatement.solve_batch(b.n, b.p, b.m, *b.fields(), max_iter=max_iter or default_max_iter(b.p, b.m), threads=4)


@pytest.mark.parametrize("name,B", [("C2", 3000), ("C5", 3000), ("C5D", 3000)])
def test_headline_shapes(emu_lib, restatement, name, B):
    b = workloads.make(name, B)
    assert_bit_identical(lane_solve(emu_lib, b), _oracle(restatement, b), name)


@pytest.mark.parametrize("box", [False, True])
@pytest.mark.parametrize("n,p,m", [(5, 2, 12), (8, 0, 20), (8, 3, 9), (3, 1, 0), (7, 0, 0), (1, 0, 3), (2, 1, 3),
                                   (8, 8, 4), (6, 0, 64), (4, 2, 33)])
def test_random_shapes_dense_and_bound_columns(emu_lib, restatement, n, p, m, box):
    b = workloads.make_random(300, n, p, m, seed=n * 100 + m, box=box)
    assert_bit_identical(lane_solve(emu_lib, b), _oracle(restatement, b), f"R{n}_{p}_{m}")


def test_degenerate_branch_and_iteration_cap(emu_lib, restatement):
    b = _degenerate_batch(200)
    for cap in (7, 60):
        assert_bit_identical(lane_solve(emu_lib, b, cap), _oracle(restatement, b, cap), f"cap{cap}")


def test_failure_verdicts(emu_lib, restatement):
    b = workloads.make_random(64, 6, 2, 14, seed=11, box=True)
    b.CE[0::4, :, 1] = -0.5 * b.CE[0::4, :, 0]
    b.G[1::8] -= 4.0 * np.eye(6)
    b.ci0[2::8, 0] = -3.0; b.ci0[2::8, 6] = 1.0
    ref = _oracle(restatement, b)
    got = lane_solve(emu_lib, b)
    assert np.array_equal(got["status"], ref["status"])
    keep = ref["status"] != 2
    for k in ("x", "iq", "iters", "A"):
        assert np.array_equal(got[k][keep], ref[k][keep]), k
    assert np.array_equal(got["f"], ref["f"], equal_nan=True)


def test_non_finite_inputs(emu_lib, restatement):
    b = workloads.make("C2", 2048)
    b.g0[5::97, 2] = np.inf
    b.ci0[7::101, 3] = np.nan
    b.CE[11::103, 1, 0] = np.nan
    ref = _oracle(restatement, b)
    got = lane_solve(emu_lib, b)
    assert np.array_equal(got["status"], ref["status"])
    assert np.array_equal(got["x"], ref["x"], equal_nan=True) and np.array_equal(got["f"], ref["f"], equal_nan=True)


def test_non_finite_cholesky_factor(emu_lib, restatement):
    """The structural zeros of J = L^-T are only skipped while the factor is finite."""
    b = workloads.make("C2", 2048)
    b.G[3::89, 1, 4] = np.inf      # read (upper triangle): an infinite off-diagonal factor entry
    b.G[5::97, 0, 6] = np.nan
    b.G[7::101, 6, 6] = np.nan     # NaN pivot passes the `sum <= 0` test of the reference
    b.G[11::103, 2, 2] = np.inf
    b.G[13::107, 0, 3] = 1e200     # overflows further down the factorisation
    ref = _oracle(restatement, b)
    got = lane_solve(emu_lib, b)
    assert np.array_equal(got["status"], ref["status"])
    assert np.array_equal(got["x"], ref["x"], equal_nan=True) and np.array_equal(got["f"], ref["f"], equal_nan=True)
    assert np.array_equal(got["iters"], ref["iters"])


@pytest.mark.parametrize("name,B", [("C2", 3000), ("C5", 3000)])
def test_column_compressed_constraints(emu_lib, restatement, name, B):
    """mgqp_batch::CI_val / CI_row: the solver never sees the dense block, results stay bit-identical."""
    b = workloads.make(name, B)
    b.ci0[3::41, 5] = np.nan            # non-finite data takes the full-sum scan path
    b.g0[4::43, 1] = np.inf
    b.CI[5::47, :, 9] = 0.0             # an all-zero column
    b.CI[6::53, 2, 2] = -0.0
    assert_bit_identical(lane_solve(emu_lib, b, compressed=_compress(b.CI)), _oracle(restatement, b), name)


def test_column_compressed_odd_shapes(emu_lib, restatement):
    for n, p, m in [(5, 1, 11), (8, 2, 16), (3, 0, 6), (7, 1, 63)]:
        b = workloads.make_random(200, n, p, m, seed=7 * n + m)
        rng = np.random.default_rng(m)
        rows = rng.integers(0, n, size=(b.batch, m))
        vals = rng.choice([-1.0, 1.0, 0.5, -2.5], size=(b.batch, m))
        b.CI[...] = 0.0
        np.put_along_axis(b.CI, rows[:, None, :], vals[:, None, :], axis=1)
        assert_bit_identical(lane_solve(emu_lib, b, compressed=_compress(b.CI)), _oracle(restatement, b), f"R{n}_{p}_{m}")


def test_dense_input_unit_nonunit_and_signed_zero_columns(emu_lib, restatement):
    """Phase B sorts the dense CI block into +-e_k bytes, empty columns and exact-path (GEN) columns:
    non-unit bounds, two-nonzero columns and -0.0 entries must all come out bit-identical."""
    for n, p, m in [(7, 1, 42), (5, 0, 13), (8, 2, 64), (2, 0, 40), (7, 7, 3)]:
        b = workloads.make_random(257, n, p, m, seed=31 * n + m)
        rng = np.random.default_rng(n + m)
        rows = rng.integers(0, n, size=(b.batch, m))
        vals = rng.choice([-1.0, 1.0, 1.0, -1.0, 0.5, -2.5, 0.0], size=(b.batch, m))
        b.CI[...] = 0.0
        np.put_along_axis(b.CI, rows[:, None, :], vals[:, None, :], axis=1)
        b.CI[3::7, 0, m // 2] = -0.0                      # a -0.0 entry sends its column down the exact path
        b.CI[5::11, n - 1, 0] += 0.25                     # a second nonzero / a non-unit value
        b.CI[::2, :, :] = np.where(b.CI[::2] == 0.5, 1.0, b.CI[::2])   # every other QP: units and zeros only
        assert_bit_identical(lane_solve(emu_lib, b), _oracle(restatement, b), f"U{n}_{p}_{m}")


def test_scan_ties_keep_the_first_index(emu_lib, restatement):
    """Duplicated bound columns give exactly equal s_i: the reference keeps the first index of the minimum."""
    b = workloads.make("C2", 1500)
    for src, dst in [(5, 20), (5, 33), (1, 2), (40, 3), (14, 41)]:
        b.CI[:, :, dst] = b.CI[:, :, src]
        b.ci0[:, dst] = b.ci0[:, src]
    b.ci0[:, [5, 20, 33, 1, 2]] -= 1.5      # make the duplicated bounds bite
    assert_bit_identical(lane_solve(emu_lib, b), _oracle(restatement, b), "ties")


def test_sparse_direction_takes_the_sweeps_fallback(emu_lib, restatement):
    """G = I with bound constraints: d = J' n has exact zeros, which the split division of the add_constraint sweep
    hands back to the general routine (the emulator applies the device's range tests): both paths, same bits."""
    rng = np.random.default_rng(77)
    B, n, m = 3001, 7, 14
    G = np.broadcast_to(np.eye(n), (B, n, n)).copy()
    G[::3] *= rng.uniform(0.5, 2.0, (len(G[::3]), 1, 1))
    g0 = rng.standard_normal((B, n)) * 3
    CE = np.zeros((B, n, 1)); CE[:, 2, 0] = 1.0
    CE[::5, 4, 0] = -0.5
    ce0 = rng.standard_normal((B, 1)) * 0.1
    CI = np.zeros((B, n, m))
    CI[:, np.arange(n), np.arange(n)] = 1.0
    CI[:, np.arange(n), n + np.arange(n)] = -1.0
    ci0 = np.abs(rng.standard_normal((B, m))) * 0.7 + 0.2
    b = workloads.QPBatch("sparse7", n, 1, m, G, g0, CE, ce0, CI, ci0)
    ref = _oracle(restatement, b)
    assert ref["iters"].mean() > 2
    assert_bit_identical(lane_solve(emu_lib, b), ref, "sparse7")
    # tiny and huge scales: operands outside the inline path's range
    for scale in (1e-160, 1e150):
        b2 = workloads.QPBatch("scaled", n, 1, m, G, g0 * scale, CE, ce0 * scale, CI, ci0 * scale)
        assert_bit_identical(lane_solve(emu_lib, b2), _oracle(restatement, b2), f"scaled{scale}")
