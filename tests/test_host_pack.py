"""CPU: the host side of the column-compressed constraint transport (mgqp_pack_bound_columns in
include/mgqp_b200.h) -- a codec, no QP arithmetic: decoding its output must give back the dense CI
block, and anything that is not a bound on one variable must be refused."""
import numpy as np
import pytest

from mgqp_b200 import workloads


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    import mgqp_b200
    return mgqp_b200


def _decode(val, row, n):
    B, m = val.shape
    CI = np.zeros((B, n, m))
    np.put_along_axis(CI, row[:, None, :].astype(np.int64), val[:, None, :], axis=1)
    return CI


@pytest.mark.parametrize("name", ["C2", "C5"])
@pytest.mark.parametrize("threads", [1, 0])
def test_round_trip_on_the_headline_shapes(lib, name, threads):
    b = workloads.make(name, 5003)
    out = lib.pack_bound_columns(b.CI, threads=threads)
    assert out is not None
    val, row = out
    assert val.shape == (b.batch, b.m) and row.dtype == np.uint8 and (row < b.n).all()
    assert np.array_equal(_decode(val, row, b.n), b.CI)


def test_random_bound_columns_zero_columns_signed_zeros_and_nan(lib):
    rng = np.random.default_rng(5)
    for n, m in [(1, 1), (3, 6), (5, 11), (7, 42), (8, 64), (8, 63)]:
        B = 257
        rows = rng.integers(0, n, size=(B, m))
        vals = rng.standard_normal((B, m))
        CI = np.zeros((B, n, m))
        np.put_along_axis(CI, rows[:, None, :], vals[:, None, :], axis=1)
        CI[::7, :, 0] = 0.0                      # empty column
        CI[1::7, :, m - 1] = 0.0
        CI[1::7, n - 1, m - 1] = np.nan          # NaN is a nonzero
        CI[2::7][CI[2::7] == 0] = -0.0           # negative zeros stay zeros
        val, row = lib.pack_bound_columns(CI)
        dec = _decode(val, row, n)
        assert np.array_equal(dec, CI, equal_nan=True), (n, m)
        assert (val[::7, 0] == 0).all()


def test_refuses_columns_with_two_nonzeros(lib):
    b = workloads.make("C2", 4096)
    assert lib.pack_bound_columns(b.CI) is not None
    b.CI[4095, 4, 17] = 1e-300                   # one extra entry in the very last QP
    assert b.CI[4095, :, 17].astype(bool).sum() == 2
    assert lib.pack_bound_columns(b.CI) is None
    assert lib.pack_bound_columns(workloads.make_random(100, 6, 1, 12, seed=3).CI) is None


def test_empty_and_degenerate_sizes(lib):
    assert lib.pack_bound_columns(np.zeros((0, 7, 42))) is not None
    val, row = lib.pack_bound_columns(np.zeros((5, 7, 0)))
    assert val.shape == (5, 0)
    assert lib.pack_bound_columns(np.zeros((3, 7, 65))) is None      # beyond the lane kernel's m <= 64


def test_pack_thread_count_follows_the_environment(lib, monkeypatch):
    """The host pipeline decides between packed and dense transport from this number (one rank per GPU
    shares the box's cores: bench.py hands every rank cores / local world size)."""
    L = lib.lib()
    L.mgqp_host_pack_threads.restype = int
    monkeypatch.setenv("MGQP_HOST_THREADS", "5")
    assert L.mgqp_host_pack_threads() == 5
    monkeypatch.setenv("MGQP_HOST_THREADS", "1000")
    assert L.mgqp_host_pack_threads() == 64
    monkeypatch.delenv("MGQP_HOST_THREADS")
    import os
    assert 1 <= L.mgqp_host_pack_threads() <= max(1, len(os.sched_getaffinity(0)))


def test_blocks_all_equal(lib):
    import ctypes as C
    L = lib.lib()
    L.mgqp_blocks_all_equal.restype = C.c_int
    L.mgqp_blocks_all_equal.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int]
    b = workloads.make("S2", 20000)       # the shipped shape: G = I, g0 = 0 in every QP
    for th in (1, 0):
        assert L.mgqp_blocks_all_equal(b.G.ctypes.data, 196, 196, b.batch, th) == 1
        assert L.mgqp_blocks_all_equal(b.g0.ctypes.data, 14, 14, b.batch, th) == 1
    b.G[19999, 13, 13] = np.nextafter(1.0, 2.0)
    assert L.mgqp_blocks_all_equal(b.G.ctypes.data, 196, 196, b.batch, 0) == 0
    b.g0[7, 0] = -0.0                       # bitwise, not numeric, equality
    assert L.mgqp_blocks_all_equal(b.g0.ctypes.data, 14, 14, b.batch, 0) == 0
    c = workloads.make("C2", 5000)
    assert L.mgqp_blocks_all_equal(c.G.ctypes.data, 49, 49, c.batch, 0) == 0
    assert L.mgqp_blocks_all_equal(c.G.ctypes.data, 49, 49, 1, 0) == 1


_ISA_SNIPPET = r"""
import sys, hashlib, numpy as np
sys.path.insert(0, %r)
import mgqp_b200
rng = np.random.default_rng(11)
h = hashlib.sha256()
for n, m in [(1, 1), (3, 6), (5, 11), (7, 42), (8, 64), (8, 63), (7, 9)]:
    B = 301
    rows = rng.integers(0, n, size=(B, m))
    vals = rng.standard_normal((B, m))
    CI = np.zeros((B, n, m))
    np.put_along_axis(CI, rows[:, None, :], vals[:, None, :], axis=1)
    CI[::7, :, 0] = 0.0
    CI[1::7, :, m - 1] = 0.0
    CI[1::7, n - 1, m - 1] = np.nan
    CI[2::7][CI[2::7] == 0] = -0.0
    val, row = mgqp_b200.pack_bound_columns(CI)
    dec = np.zeros((B, n, m))
    np.put_along_axis(dec, row[:, None, :].astype(np.int64), val[:, None, :], axis=1)
    assert np.array_equal(dec, CI, equal_nan=True), (n, m)
    h.update(val.tobytes()); h.update(row.tobytes())
    CI[B - 1, 0, m - 1] = 1.0
    CI[B - 1, n - 1, m - 1] = 2.0
    if n > 1:
        assert mgqp_b200.pack_bound_columns(CI) is None, (n, m)
for n, m in [(1, 1), (3, 6), (7, 42), (8, 64), (8, 63), (14, 28)]:
    B = 203
    rows = rng.integers(0, n, size=(B, m))
    sgn = np.where(rng.integers(0, 2, size=(B, m)) == 1, 1.0, -1.0)
    CI = np.zeros((B, n, m))
    np.put_along_axis(CI, rows[:, None, :], sgn[:, None, :], axis=1)
    CI[::5, :, 0] = 0.0
    CI[1::5][CI[1::5] == 0] = -0.0
    code = mgqp_b200.pack_unit_columns(CI)
    assert code is not None, (n, m)
    empty = (code & 0x40) != 0
    assert (empty[::5, 0]).all() and empty.sum() == len(CI[::5])
    val = np.where(empty, 0.0, np.where(code & 0x80, -1.0, 1.0))
    dec = np.zeros((B, n, m))
    np.put_along_axis(dec, np.where(empty, 0, code & 0x3f)[:, None, :].astype(np.int64), val[:, None, :], axis=1)
    assert np.array_equal(dec, CI), (n, m)
    h.update(code.tobytes())
    bad = CI.copy(); bad[B - 1, n - 1, m - 1] = 0.5 if CI[B - 1, n - 1, m - 1] == 0 or True else 0
    bad[B - 1, :, m - 1] = 0.0; bad[B - 1, n - 1, m - 1] = 0.5          # a non-unit coefficient
    assert mgqp_b200.pack_unit_columns(bad) is None, (n, m)
    bad = CI.copy(); bad[0, :, 0] = 0.0; bad[0, 0, 0] = np.nan          # NaN is not a unit
    assert mgqp_b200.pack_unit_columns(bad) is None, (n, m)
    if n > 1:
        bad = CI.copy(); bad[B // 2, :, m // 2] = 0.0; bad[B // 2, 0, m // 2] = 1.0; bad[B // 2, n - 1, m // 2] = -1.0
        assert mgqp_b200.pack_unit_columns(bad) is None, (n, m)
print(h.hexdigest())
"""


def test_every_instruction_set_of_the_packer_gives_the_same_bytes(lib):
    """generic / AVX2 / AVX-512 (picked once per process: one subprocess each; an ISA the CPU lacks falls back)"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digests = {}
    for isa in ("generic", "avx2", "avx512"):
        env = dict(os.environ, MGQP_HOST_PACK_ISA=isa)
        r = subprocess.run([sys.executable, "-c", _ISA_SNIPPET % root], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        digests[isa] = r.stdout.strip().splitlines()[-1]
    assert len(set(digests.values())) == 1, digests


def test_pool_grows_when_the_thread_budget_is_raised(lib, monkeypatch):
    """A torchrun rank packs with its share of the cores first and may drive every device later (bench.py's
    multi-device record): the pool must follow MGQP_HOST_THREADS upwards."""
    b = workloads.make("C2", 20000)
    monkeypatch.setenv("MGQP_HOST_THREADS", "2")
    first = lib.pack_bound_columns(b.CI)
    monkeypatch.setenv("MGQP_HOST_THREADS", "7")
    second = lib.pack_bound_columns(b.CI)
    third = lib.pack_bound_columns(b.CI, threads=11)
    for out in (second, third):
        assert np.array_equal(out[0], first[0]) and np.array_equal(out[1], first[1])
