"""CPU: the DEVICE solver body (csrc/gi_small.cuh) compiled for the host and run under a fiber warp
emulator (tests/emu) must be bit-identical to the oracle -- x, f, verdict, active set, step count.
This checks the lock-step control flow (predicated dual/full/partial/degenerate branches, warp
primitives in uniform control flow) without a GPU; the GPU tests then only have to establish that
the hardware executes the same program."""
import numpy as np
import pytest

from conftest import default_max_iter
from helpers import assert_bit_identical, assert_parity, emu_solve
from mgqp_b200 import workloads


def _oracle(restatement, b, max_iter=0):
    return restatement.solve_batch(b.n, b.p, b.m, *b.fields(), max_iter=max_iter or default_max_iter(b.p, b.m), threads=4)


@pytest.mark.parametrize("name,B,gw", [("C2", 403, 8), ("C5", 402, 8), ("C3", 203, 16), ("S0", 201, 16),
                                       ("S2", 201, 16), ("C2", 61, 16), ("C2", 33, 32)])
def test_benchmark_shapes(emu_lib, restatement, name, B, gw):
    b = workloads.make(name, B)
    assert_bit_identical(emu_solve(emu_lib, b, gw), _oracle(restatement, b), f"{name}/gw{gw}")


@pytest.mark.parametrize("n,p,m,gw", [(5, 2, 12, 8), (8, 0, 20, 8), (8, 3, 9, 8), (3, 1, 0, 8), (7, 0, 0, 8),
                                      (12, 4, 30, 16), (16, 2, 40, 16), (20, 5, 50, 32), (32, 4, 70, 32),
                                      (1, 0, 3, 8), (2, 1, 3, 8), (8, 8, 4, 8), (6, 0, 200, 8)])
def test_random_shapes(emu_lib, restatement, n, p, m, gw):
    b = workloads.make_random(67, n, p, m, seed=n * 100 + m)
    assert_bit_identical(emu_solve(emu_lib, b, gw), _oracle(restatement, b), f"R{n}_{p}_{m}")


def test_failure_verdicts(emu_lib, restatement):
    """Not-PD G, dependent equalities and infeasible bounds mixed into one batch (each group of
    lanes must retire on its own verdict while its warp-mates keep iterating)."""
    b = workloads.make_random(64, 6, 2, 14, seed=11, box=True)
    for i in range(0, 64, 4):
        b.CE[i, :, 1] = -0.5 * b.CE[i, :, 0]      # status 3
    for i in range(1, 64, 8):
        b.G[i] -= 4.0 * np.eye(6)                  # status 2
    for i in range(2, 64, 8):
        b.ci0[i, 0] = -3.0; b.ci0[i, 6] = 1.0      # x0 >= 3 and x0 <= 1: status 1
    ref = _oracle(restatement, b)
    assert set(np.unique(ref["status"])) >= {0, 1, 2, 3}
    got = emu_solve(emu_lib, b, 8)
    assert np.array_equal(got["status"], ref["status"])
    keep = ref["status"] != 2   # after a failed Cholesky the reference has thrown: x, iq are not defined
    for k in ("x", "iq", "iters", "A"):
        assert np.array_equal(got[k][keep], ref[k][keep]), k
    assert np.array_equal(got["f"], ref["f"], equal_nan=True)  # status 2: f carries the bad pivot


def _degenerate_batch(B=96):
    """Badly scaled constraint normals: once a 1e10-scaled bound is active, R_norm ~ 1e10 and adding a
    1e-7-scaled one fails add_constraint's |d[iq]| <= eps*R_norm test while z.z > eps still holds, i.e.
    the reference's degenerate branch (note N3: it then usually spins forever)."""
    b = workloads.make_random(B, 5, 1, 12, seed=3)
    b.CI[:, :, 0:6] *= 1e10; b.ci0[:, 0:6] *= 1e10
    b.CI[:, :, 6:12] *= 1e-7; b.ci0[:, 6:12] *= 1e-7
    b.g0 *= 4.0
    return b


def test_degenerate_branch_and_iteration_cap(emu_lib, restatement):
    import ctypes as C
    b = _degenerate_batch()
    cnt = C.c_longlong.in_dll(restatement.lib, "gi_oracle_degenerate_count")
    for cap in (7, 60):
        cnt.value = 0
        ref = restatement.solve_batch(b.n, b.p, b.m, *b.fields(), max_iter=cap, threads=1)
        assert cnt.value > 50, "the batch is meant to exercise the degenerate-add branch"
        got = emu_solve(emu_lib, b, 8, max_iter=cap)
        assert_bit_identical(got, ref, f"cap{cap}")
    assert (ref["status"] == 4).any() and (ref["status"] == 0).any()


def _identity_mix(B):
    """As-shipped S0 batch (G = I) with a few QPs whose G is only ALMOST the identity: -0.0 off the
    diagonal, a diagonal entry one ulp above 1, a genuinely different SPD matrix."""
    b = workloads.make("S0", B)
    b.G[3, 2, 5] = -0.0
    b.G[3, 5, 2] = -0.0
    b.G[10, 4, 4] = np.nextafter(1.0, 2.0)
    b.G[17] = workloads.make_random(1, 14, 0, 0, seed=5).G[0]
    return b


def test_identity_fast_path_and_its_fallback(emu_lib, restatement):
    b = _identity_mix(64)
    assert_bit_identical(emu_solve(emu_lib, b, 16), _oracle(restatement, b), "S0 identity mix")
    b = workloads.make("S2", 64)          # every warp all-identity: the fast path itself
    assert_bit_identical(emu_solve(emu_lib, b, 16), _oracle(restatement, b), "S2")


def test_fast_math_mode_within_tolerance(emu_lib, restatement):
    """FMA + tree reductions: not bit-identical by design, but inside the 1e-9 gate."""
    b = workloads.make("C2", 301)
    got = emu_solve(emu_lib, b, 8, fma=1)
    assert_parity(got, _oracle(restatement, b), "C2/fast")


def _emu_solve_comp(lib, b, gw, fma=0):
    """The group kernel's solver with the inequality block kept column-compressed in shared memory (values | rows)."""
    import ctypes as C
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    B, n, p, m = b.batch, b.n, b.p, b.m
    nz = b.CI != 0
    assert (nz.sum(axis=1) <= 1).all()
    row = nz.argmax(axis=1).astype(np.uint8)
    val = np.take_along_axis(b.CI, row[:, None, :].astype(np.int64), axis=1)[:, 0, :]
    val = np.where(nz.any(axis=1), val, 0.0)
    G, g0, CE, ce0, _, ci0 = [np.ascontiguousarray(a, dtype=np.float64) for a in b.fields()]
    val = np.ascontiguousarray(val)
    x = np.zeros((B, n)); f = np.zeros(B)
    st = np.zeros(B, np.int32); iq = np.zeros(B, np.int32); A = np.zeros((B, p + m), np.int32); it = np.zeros(B, np.int32)
    lib.emu_solve_batch_comp.restype = C.c_int
    rc = lib.emu_solve_batch_comp(C.c_int(n), C.c_int(p), C.c_int(m), C.c_longlong(B), G.ctypes.data_as(dp),
                                  g0.ctypes.data_as(dp), CE.ctypes.data_as(dp), ce0.ctypes.data_as(dp),
                                  val.ctypes.data_as(dp), row.ctypes.data_as(C.POINTER(C.c_ubyte)), ci0.ctypes.data_as(dp),
                                  x.ctypes.data_as(dp), f.ctypes.data_as(dp), *[a.ctypes.data_as(ip) for a in (st, iq, A, it)],
                                  C.c_int(0), C.c_int(gw), C.c_int(fma))
    assert rc == 0
    return dict(x=x, f=f, status=st, iq=iq, A=A, iters=it)


@pytest.mark.parametrize("name,B,gw", [("S0", 301, 16), ("S2", 301, 16), ("C2", 203, 8), ("C5", 203, 8), ("C5D", 150, 16)])
def test_compressed_constraints_stay_compressed_in_shared_memory(emu_lib, restatement, name, B, gw):
    b = workloads.make(name, B)
    if name in ("S0", "S2"):
        b.ci0[::7, 3] = -0.3                 # bounds that bite: iterations past the equality phase
        b.ci0[::7, 3 + 14] = -0.3            # ... and contradict each other (infeasible verdicts)
        b.ci0[1::7, 5] = -0.05
        b.g0[2::11] = np.random.default_rng(1).standard_normal((len(b.g0[2::11]), b.n))
    b.CI[3::13, :, 1] = 0.0                  # an empty column
    b.CI[4::13][b.CI[4::13] == 0] = -0.0     # negative zeros stay zeros
    b.CI[5::13, :, 2] *= 0.37                # a non-unit coefficient
    ref = _oracle(restatement, b)
    assert ref["iters"].max() > 1
    assert_bit_identical(_emu_solve_comp(emu_lib, b, gw), ref, name + " compressed")


def test_compressed_constraints_with_non_finite_iterates(emu_lib, restatement):
    """0 * inf = NaN: a non-finite x entry poisons every column's sum except through the column's own row."""
    b = workloads.make("S2", 120)
    b.g0[0::6, 2] = np.inf
    b.g0[1::6, 4] = -np.inf
    b.g0[2::6, 1] = np.nan
    b.g0[3::6, [0, 7]] = np.inf
    b.ce0[4::6, 0] = np.inf
    ref = _oracle(restatement, b)
    assert_bit_identical(_emu_solve_comp(emu_lib, b, 16), ref, "S2 compressed, non-finite")
    assert_bit_identical(emu_solve(emu_lib, b, 16), ref, "S2 dense, non-finite")
