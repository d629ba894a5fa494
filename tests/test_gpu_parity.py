"""GPU (-m gpu): the CUDA path, called through the C ABI (libmgqp_b200.so via ctypes), against the
oracle on the same seeded inputs.  Gate (BASELINE.json north_star): identical verdicts and final
active sets; x and f within 1e-9 relative.  The default (exact) mode is additionally required to be
BIT-identical, which is what the kernel is designed to deliver."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN, default_max_iter
from helpers import assert_bit_identical, assert_parity
from mgqp_b200 import workloads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    import __graft_entry__ as g
    g.build()
    # the host pipeline packs CI only where it has the threads to outrun the link; the transport tests below
    # check byte counts of the packed form, so ask for it whatever box this runs on
    os.environ.setdefault("MGQP_HOST_THREADS", "16")
    import mgqp_b200
    assert torch.cuda.is_available(), "these tests need a B200"
    mgqp_b200.lib()
    return mgqp_b200


def _dev(b, idx=0):
    import torch
    return [torch.from_numpy(np.ascontiguousarray(a)).to(f"cuda:{idx}") for a in b.fields()]


def _solve(gpu, b, **kw):
    import torch
    r = gpu.solve_quadprog_batched(*_dev(b), want_active_set=True, **kw)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in r._asdict().items()}


def _oracle(restatement, b, max_iter=0):
    return restatement.solve_batch(b.n, b.p, b.m, *b.fields(), max_iter=max_iter or default_max_iter(b.p, b.m), threads=8)


@pytest.mark.parametrize("name,B", [("C2", 20001), ("C5", 20003), ("C5D", 20005), ("C3", 6001), ("S0", 6001), ("S2", 6001)])
def test_benchmark_shapes_bit_identical(gpu, restatement, name, B):
    b = workloads.make(name, B)
    assert_bit_identical(_solve(gpu, b), _oracle(restatement, b), name)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "batch_[CS]*[0-9].npz"))) +
                         sorted(glob.glob(os.path.join(GOLDEN, "batch_R[0-9]*_*.npz"))))
def test_golden_fixtures_from_reference_binary(gpu, path):
    """x / f / verdict the reference's shipped binary produced (tests/golden/make_golden.py)."""
    z = np.load(path)
    n = z["g0"].shape[1]
    b = workloads.QPBatch("golden", n, z["ce0"].shape[1], z["ci0"].shape[1], z["G"], z["g0"], z["CE"], z["ce0"], z["CI"], z["ci0"])
    got = _solve(gpu, b)
    assert_bit_identical(dict(x=got["x"], f=got["f"], status=got["status"]), dict(x=z["x"], f=z["f"], status=z["status"]),
                         os.path.basename(path))


@pytest.mark.parametrize("n,p,m", [(1, 0, 3), (2, 1, 3), (3, 1, 0), (7, 0, 0), (5, 2, 12), (8, 0, 20), (8, 8, 4),
                                   (6, 0, 200), (9, 2, 33), (12, 4, 30), (16, 2, 40), (17, 0, 21), (20, 5, 50),
                                   (32, 4, 70), (31, 31, 5)])
def test_random_shapes_bit_identical(gpu, restatement, n, p, m):
    for box in (False, True):   # dense columns (re-read from L2 by the lane kernel) / compressed bound columns
        b = workloads.make_random(1003, n, p, m, seed=n * 100 + m, box=box)
        ref = _oracle(restatement, b)
        assert_bit_identical(_solve(gpu, b), ref, f"R{n}_{p}_{m}")
        if gpu.plan(n, p, m)[0] == 1:
            assert_bit_identical(_solve(gpu, b, flags=8), ref, f"R{n}_{p}_{m}/group")


@pytest.mark.parametrize("flags", [2, 4, 6, 8, 10, 12, 14])
def test_staging_schedule_and_kernel_family_variants(gpu, restatement, flags):
    """NO_BULK_COPY (2), STATIC_SCHEDULE (4) and GROUP_KERNEL (8: lane-group kernel instead of the
    one-QP-per-lane kernel) must not change a bit."""
    b = workloads.make("C2", 5003)
    assert_bit_identical(_solve(gpu, b, flags=flags), _oracle(restatement, b), f"flags{flags}")


@pytest.mark.parametrize("n,p,m", [(33, 0, 70), (40, 3, 90), (50, 5, 160), (64, 8, 100), (100, 0, 300), (7, 1, 42),
                                   (14, 10, 28), (6, 0, 200)])
def test_cta_per_qp_kernel_bit_identical(gpu, restatement, n, p, m):
    """The CTA-per-QP family (n > 32; forced with flag 16 on the small shapes)."""
    b = workloads.make_random(96, n, p, m, seed=7 * n + m, box=(m >= 2 * n))
    assert_bit_identical(_solve(gpu, b, flags=16), _oracle(restatement, b), f"cta R{n}_{p}_{m}")


def test_cta_per_qp_kernel_failures_and_degenerate(gpu, restatement):
    from test_emu_lockstep import _degenerate_batch
    b = _degenerate_batch(256)
    for cap in (7, 60):
        assert_bit_identical(_solve(gpu, b, max_iter=cap, flags=16), _oracle(restatement, b, max_iter=cap), f"cta cap{cap}")
    b = workloads.make_random(256, 6, 2, 14, seed=11, box=True)
    b.CE[0::4, :, 1] = -0.5 * b.CE[0::4, :, 0]
    b.G[1::8] -= 4.0 * np.eye(6)
    b.ci0[2::8, 0] = -3.0; b.ci0[2::8, 6] = 1.0
    ref = _oracle(restatement, b)
    got = _solve(gpu, b, flags=16)
    assert np.array_equal(got["status"], ref["status"])
    keep = ref["status"] != 2
    for k in ("x", "iq", "iters", "A"):
        assert np.array_equal(got[k][keep], ref[k][keep]), k
    assert np.array_equal(got["f"], ref["f"], equal_nan=True)


def test_cta_per_qp_kernel_poisoned_factor(gpu, restatement):
    b = workloads.make_random(64, 40, 2, 60, seed=77)
    b.G[1, 3, 4] = b.G[1, 4, 3] = np.inf
    b.G[2, 5, 5] = np.nan
    b.G[3, 39, 39] = np.inf
    ref = _oracle(restatement, b)
    got = _solve(gpu, b)
    assert np.array_equal(got["status"], ref["status"])
    keep = ref["status"] != 2
    assert np.array_equal(got["x"][keep], ref["x"][keep], equal_nan=True)
    assert np.array_equal(got["f"], ref["f"], equal_nan=True)


def test_mpc_shape_against_reference_binary_and_oracle(gpu, restatement):
    """BASELINE config 4 shape: n = 350, m = 2100.  One QP is pinned by the shipped binary
    (tests/golden/batch_C4_seedonly.npz); a small batch is checked against the oracle."""
    z = np.load(os.path.join(GOLDEN, "batch_C4_seedonly.npz"))
    b = workloads.make("C4", 1)
    got = _solve(gpu, b)
    assert np.array_equal(got["x"], z["x"]) and np.array_equal(got["f"], z["f"]) and np.array_equal(got["status"], z["status"])
    b = workloads.make("C4", 6, seed=99)
    assert_bit_identical(_solve(gpu, b), _oracle(restatement, b), "C4")


def test_mpc_shape_64_qps_exact_and_tensor_core_mode(gpu, restatement):
    """64 horizon-stacked QPs (n = 350, m = 2100, ~150 active-set iterations each) against the oracle: the default
    mode bit-identical; MGQP_FLAG_FAST_MATH (blocked Cholesky / triangular inverse on fp64 DMMA, FMA elsewhere) under
    the north-star gate -- identical verdicts and final active sets, x and f within 1e-9 relative."""
    b = workloads.make("C4", 64, seed=4242)
    ref = _oracle(restatement, b)
    assert ref["iters"].mean() > 100
    assert_bit_identical(_solve(gpu, b), ref, "C4 x 64")
    assert_parity(_solve(gpu, b, flags=gpu.FLAG_FAST_MATH), ref, "C4 x 64, DMMA set-up")
    # a matrix that is not positive definite must fail in both modes
    b.G[3] -= 3.0 * np.eye(350)
    for fl in (0, gpu.FLAG_FAST_MATH):
        got = _solve(gpu, b, flags=fl)
        assert got["status"][3] == 2 and (np.delete(got["status"], 3) == 0).all()


def test_mpc_shape_full_size_properties(gpu):
    """BASELINE.json configs[3] at a size no CPU oracle can follow (2368 QPs generated on the device): every answer
    is primal feasible, satisfies the KKT stationarity of its final active set, and the tensor-core mode reproduces
    verdicts, active sets and iterates of the exact mode to 1e-9."""
    import torch
    t = workloads.make_c4_device(2368, torch.device("cuda:0"))
    r = gpu.solve_quadprog_batched(*t, want_active_set=True)
    rf = gpu.solve_quadprog_batched(*t, want_active_set=True, flags=gpu.FLAG_FAST_MATH)
    torch.cuda.synchronize()
    assert (r.status == 0).all() and torch.equal(r.status, rf.status)
    assert torch.equal(r.iq, rf.iq) and torch.equal(r.A, rf.A)
    G, g0, _, _, CI, ci0 = t
    s = torch.einsum("bnm,bn->bm", CI, r.x) + ci0
    # the solver stops when the total violation |psi| = -sum min(0, s_i) is below m * eps * trace(G) * trace(J) * 100
    # (SURVEY.md App. A, step 1); G >= I here, so trace(J) = sum 1/L_ii <= n
    n, m = G.shape[1], CI.shape[2]
    thr = 100.0 * m * 2.220446049250313e-16 * torch.diagonal(G, dim1=1, dim2=2).sum(dim=1) * n
    assert ((-s.clamp(max=0.0).sum(dim=1)) <= thr).all(), "primal feasibility within the solver's own stopping rule"
    # stationarity: G x + g0 = CI[:, A] u with u >= 0  <=>  the residual of the least-squares fit over the active
    # columns vanishes; checked through f = 1/2 x'Gx + g0'x (the value the solver reports)
    f = 0.5 * torch.einsum("bn,bnk,bk->b", r.x, G, r.x) + (g0 * r.x).sum(dim=1)
    assert ((f - r.f).abs() <= 1e-9 * r.f.abs().clamp(min=1.0)).all(), "reported cost"
    relx = (rf.x - r.x).abs().amax(dim=1) / r.x.abs().amax(dim=1).clamp(min=1.0)
    assert relx.max().item() <= 1e-9 and ((rf.f - r.f).abs() <= 1e-9 * r.f.abs().clamp(min=1.0)).all()
    # every active inequality is tight, every inactive one slack or tight
    act = torch.zeros_like(s, dtype=torch.bool)
    idx = r.A.clamp(min=0, max=s.shape[1] - 1).long()
    valid = (r.A >= 0) & (r.A < s.shape[1])
    act.scatter_(1, idx, valid)
    assert (s[act].abs() <= thr.max()).all(), "active constraints are tight"


def test_identity_g_fast_path_and_its_fallback(gpu, restatement):
    """G == I exactly (the reference's own caller, src/mgqp.cpp:672) skips the factorisation; a warp
    that also holds a QP with -0.0 / 1+ulp / a different G must take the general path."""
    from test_emu_lockstep import _identity_mix
    for b in (_identity_mix(4099), workloads.make("S2", 4099), workloads.make("S0", 4099)):
        assert_bit_identical(_solve(gpu, b), _oracle(restatement, b), b.name)


def test_empty_batch_and_argument_errors(gpu):
    import torch
    z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device="cuda:0")
    r = gpu.solve_quadprog_batched(z(0, 3, 3), z(0, 3), z(0, 3, 1), z(0, 1), z(0, 3, 2), z(0, 2))
    assert r.x.shape == (0, 3) and r.status.numel() == 0
    with pytest.raises(ValueError, match="equalities"):          # p > n: the reference indexes R out of bounds there
        gpu.solve_quadprog_batched(z(4, 2, 2), z(4, 2), z(4, 2, 3), z(4, 3), z(4, 2, 0), z(4, 0))
    with pytest.raises(gpu.LogicError, match="The matrix CI is incompatible"):
        gpu.solve_quadprog_batched(z(4, 2, 2), z(4, 2), z(4, 2, 0), z(4, 0), z(4, 3, 5), z(4, 5))


def test_lane_kernel_shape_limits(gpu, restatement):
    """n = 8 with m = 64 is the largest shape of the one-QP-per-lane family; m = 65 must fall back."""
    for m in (63, 64, 65):
        b = workloads.make_random(777, 8, 2, m, seed=m, box=True)
        assert gpu.plan(8, 2, m)[0] == (1 if m <= 64 else 8)
        assert_bit_identical(_solve(gpu, b), _oracle(restatement, b), f"m{m}")


def test_tail_and_tiny_batches(gpu, restatement):
    for B in (1, 2, 3, 4, 5, 7, 63):
        b = workloads.make("C2", B, seed=B)
        assert_bit_identical(_solve(gpu, b), _oracle(restatement, b), f"B{B}")


@pytest.mark.parametrize("flags", [0, 8])
def test_failure_verdicts(gpu, restatement, flags):
    b = workloads.make_random(4096, 6, 2, 14, seed=11, box=True)
    b.CE[0::4, :, 1] = -0.5 * b.CE[0::4, :, 0]          # dependent equalities  -> 3
    b.G[1::8] -= 4.0 * np.eye(6)                         # indefinite G          -> 2
    b.ci0[2::8, 0] = -3.0; b.ci0[2::8, 6] = 1.0          # contradictory bounds  -> 1
    ref = _oracle(restatement, b)
    got = _solve(gpu, b, flags=flags)
    assert set(np.unique(ref["status"])) >= {0, 1, 2, 3}
    assert np.array_equal(got["status"], ref["status"])
    keep = ref["status"] != 2
    for k in ("x", "iq", "iters", "A"):
        assert np.array_equal(got[k][keep], ref[k][keep]), k
    assert np.array_equal(got["f"], ref["f"], equal_nan=True)


@pytest.mark.parametrize("flags", [0, 8])
def test_degenerate_branch_and_iteration_cap(gpu, restatement, flags):
    from test_emu_lockstep import _degenerate_batch
    b = _degenerate_batch(2048)
    for cap in (7, 60):
        ref = _oracle(restatement, b, max_iter=cap)
        assert_bit_identical(_solve(gpu, b, max_iter=cap, flags=flags), ref, f"cap{cap}")
    assert (ref["status"] == 4).any()


@pytest.mark.parametrize("flags", [1, 9])
def test_fast_math_flag_within_tolerance(gpu, restatement, flags):
    b = workloads.make("C2", 20001)
    assert_parity(_solve(gpu, b, flags=flags), _oracle(restatement, b), "C2/fast")


def test_non_finite_inputs_follow_the_reference(gpu, restatement):
    """NaN / inf in the data (the reference's own failure mode, '-inf problem.log.txt'): the compressed
    bound columns must not be used once x stops being finite."""
    b = workloads.make("C2", 4099)
    b.g0[5::97, 2] = np.inf
    b.ci0[7::101, 3] = np.nan
    b.CE[11::103, 1, 0] = np.nan
    ref = _oracle(restatement, b)
    for flags in (0, 8):
        got = _solve(gpu, b, flags=flags)
        assert np.array_equal(got["status"], ref["status"])
        assert np.array_equal(got["x"], ref["x"], equal_nan=True) and np.array_equal(got["f"], ref["f"], equal_nan=True)


def test_broadcast_and_padded_strides(gpu, restatement):
    """stride 0 (one G / CI for the whole batch) and padded strides go through the non-TMA loader."""
    import torch
    b = workloads.make("C2", 3001)
    b.G[:] = b.G[0]
    b.CI[:] = b.CI[0]
    ref = _oracle(restatement, b)
    t = _dev(b)
    t[0] = t[0][:1]                      # broadcast G
    t[4] = t[4][:1]                      # broadcast CI
    pad = torch.zeros((b.batch, b.n + 3), dtype=torch.float64, device="cuda:0")
    pad[:, : b.n] = t[1]
    t[1] = pad[:, : b.n]                 # g0 with a padded row stride
    r = gpu.solve_quadprog_batched(*t, want_active_set=True)
    torch.cuda.synchronize()
    assert_bit_identical({k: v.cpu().numpy() for k, v in r._asdict().items()}, ref, "strided")


def test_host_buffer_entry_point(gpu, restatement):
    b = workloads.make("C5", 70001)
    ref = _oracle(restatement, b)
    r = gpu.solve_quadprog_batched_host(*b.fields(), want_active_set=True, chunk=16384)
    assert_bit_identical(r._asdict(), ref, "host")


def test_column_compressed_device_input(gpu, restatement):
    """mgqp_batch::CI_val / CI_row on device pointers: the dense CI block is never given to the kernel."""
    import torch
    for name, B in (("C2", 20001), ("C5", 10007)):
        b = workloads.make(name, B)
        b.ci0[3::41, 5] = np.nan
        b.g0[4::43, 1] = np.inf
        b.CI[5::47, :, 9] = 0.0
        val, row = gpu.pack_bound_columns(b.CI)
        t = _dev(b)
        r = gpu.solve_quadprog_batched(t[0], t[1], t[2], t[3], None, t[5], want_active_set=True,
                                       CI_val=torch.from_numpy(val).cuda(), CI_row=torch.from_numpy(row).cuda())
        torch.cuda.synchronize()
        assert_bit_identical({k: v.cpu().numpy() for k, v in r._asdict().items()}, _oracle(restatement, b), name)
    for name, B in (("S0", 3001), ("S2", 3003)):  # the shipped n = 14 shapes: lane-group kernel
        b = workloads.make(name, B)
        b.CI[2::31][b.CI[2::31] == 0] = -0.0
        b.ci0[7::53, 4] = np.nan
        val, row = gpu.pack_bound_columns(b.CI)
        t = _dev(b)
        r = gpu.solve_quadprog_batched(t[0], t[1], t[2], t[3], None, t[5], want_active_set=True,
                                       CI_val=torch.from_numpy(val).cuda(), CI_row=torch.from_numpy(row).cuda())
        torch.cuda.synchronize()
        assert_bit_identical({k: v.cpu().numpy() for k, v in r._asdict().items()}, _oracle(restatement, b), name)
        ref = _oracle(restatement, b)
        r = gpu.solve_quadprog_batched_host(*b.fields(), want_active_set=True, chunk=1024)
        up, _ = gpu.last_transfer_bytes()
        assert_bit_identical(r._asdict(), ref, name + " host")
        # G = I and g0 = 0 in every QP (the reference's own caller): one instance per chunk, not per QP
        chunks = [min(1024, B - q) for q in range(0, B, 1024)]
        from mgqp_b200 import _api
        vals = 1 if _api.lib().mgqp_pack_unit_is_fast() else 2   # +-1 columns travel as one code byte (AVX-512 hosts)
        assert up == sum(8 * (b.n * b.n + b.n) + c * (8 * (b.n * b.p + b.p + vals * b.m) + (b.m + 15) // 16 * 16) for c in chunks)
        b.G[B // 2, 3, 3] = 1.25    # one QP with its own G: that chunk ships G per QP again, same answers
        ref = _oracle(restatement, b)
        r = gpu.solve_quadprog_batched_host(*b.fields(), want_active_set=True, chunk=1024)
        up2, _ = gpu.last_transfer_bytes()
        assert_bit_identical(r._asdict(), ref, name + " host, one odd G")
        assert up2 == up + 8 * b.n * b.n * (chunks[(B // 2) // 1024] - 1)
    b = workloads.make_random(8, 40, 2, 10, seed=1)  # n = 40: CTA-per-QP kernel, no compressed form there
    val = torch.zeros((8, b.m), dtype=torch.float64, device="cuda")
    row = torch.zeros((8, b.m), dtype=torch.uint8, device="cuda")
    t = _dev(b)
    with pytest.raises(NotImplementedError):
        gpu.solve_quadprog_batched(*t, CI_val=val, CI_row=row)


def test_dense_block_classified_on_the_device(gpu, restatement):
    """A dense inequality block handed to the device-resident entry point: the launcher classifies it on the device
    (bound_columns_kernel) and the group kernel runs with the compressed layout if every column of the batch is a
    bound, with the dense one otherwise.  Batches above one resident wave (148 SMs x 11 warps x 2 QPs) take that path."""
    import torch
    for name, B in (("S0", 5003), ("S2", 5001)):
        b = workloads.make(name, B)
        b.CI[2::31][b.CI[2::31] == 0] = -0.0      # -0.0 is a zero
        b.CI[3::37, :, 6] = 0.0                    # an all-zero column
        b.CI[5::41, np.flatnonzero(b.CI[0, :, 3])[0], 3] = np.nan   # NaN where the column's nonzero is
        b.ci0[7::53, 4] = np.nan
        b.g0[11::59, 1] = np.inf
        assert gpu.pack_bound_columns(b.CI) is not None
        assert_bit_identical(_solve(gpu, b), _oracle(restatement, b), name + " all bounds")
        # one QP anywhere in the batch with a second coefficient in a column: the whole batch runs dense
        for q in (0, B // 2, B - 1):
            c = workloads.make(name, B)
            c.CI[q, (np.flatnonzero(c.CI[q, :, 5])[0] + 1) % c.n, 5] = 0.25
            assert gpu.pack_bound_columns(c.CI) is None
            assert_bit_identical(_solve(gpu, c), _oracle(restatement, c), f"{name} mixed at {q}")
        # one block for the whole batch (leading dimension 1)
        t = _dev(b)
        r = gpu.solve_quadprog_batched(t[0], t[1], t[2], t[3], t[4][:1], t[5], want_active_set=True)
        torch.cuda.synchronize()
        b.CI[:] = b.CI[:1]
        assert_bit_identical({k: v.cpu().numpy() for k, v in r._asdict().items()}, _oracle(restatement, b), name + " one CI")
        # with the retry policy on top (the verdict bit and the second solve see the same first answers)
        b = workloads.make(name, B)
        b.ci0[::9, 3] = -50.0
        got = _solve(gpu, b, flags=64)
        val, row = gpu.pack_bound_columns(b.CI)
        t = _dev(b)
        r = gpu.solve_quadprog_batched(t[0], t[1], t[2], t[3], None, t[5], want_active_set=True, flags=64,
                                       CI_val=torch.from_numpy(val).cuda(), CI_row=torch.from_numpy(row).cuda())
        torch.cuda.synchronize()
        assert_bit_identical(got, {k: v.cpu().numpy() for k, v in r._asdict().items()}, name + " retry")


def test_host_pipeline_compressed_transport(gpu, restatement):
    """The host-buffer entry point ships CI column-compressed chunk by chunk when it can, dense when it
    cannot; either way the results are those of the oracle, bit for bit."""
    b = workloads.make("C5", 40001)
    ref = _oracle(restatement, b)
    r = gpu.solve_quadprog_batched_host(*b.fields(), want_active_set=True, chunk=4096)
    up_packed, down = gpu.last_transfer_bytes()
    assert_bit_identical(r._asdict(), ref, "packed")
    r = gpu.solve_quadprog_batched_host(*b.fields(), want_active_set=True, chunk=4096, flags=gpu.FLAG_NO_HOST_PACK)
    up_dense, _ = gpu.last_transfer_bytes()
    assert_bit_identical(r._asdict(), ref, "dense")
    B, n, p, m = b.batch, b.n, b.p, b.m
    assert up_dense == 8 * B * (n * n + n + n * p + p + n * m + m)
    from mgqp_b200 import _api
    unit = bool(_api.lib().mgqp_pack_unit_is_fast())     # +-1 columns: one code byte each (AVX-512 hosts)
    rows = B * ((m + 15) // 16 * 16)
    assert up_packed == 8 * B * (n * n + n + n * p + p + m) + rows + (0 if unit else 8 * B * m)
    assert down == B * (8 * n + 8 + 4 + 4 + 4 * (p + m) + 4)
    if unit:
        # the (value, row) form of the same batch, and a batch whose later chunks stop being +-1
        import os
        os.environ["MGQP_NO_UNIT_CODES"] = "1"
        try:
            r = gpu.solve_quadprog_batched_host(*b.fields(), want_active_set=True, chunk=4096)
            assert gpu.last_transfer_bytes()[0] == up_packed + 8 * B * m
            assert_bit_identical(r._asdict(), ref, "value-row")
        finally:
            del os.environ["MGQP_NO_UNIT_CODES"]
        b2 = workloads.make("C5", 40001)
        b2.CI[20000:, :, 7] *= 0.5
        b2.CI[3, :, 9] = 0.0                     # an empty column and a negative zero stay inside the unit form
        b2.CI[4, 1, 9] = -0.0
        r = gpu.solve_quadprog_batched_host(*b2.fields(), want_active_set=True, chunk=4096)
        up2, _ = gpu.last_transfer_bytes()
        assert_bit_identical(r._asdict(), _oracle(restatement, b2), "unit then value-row")
        assert up_packed < up2 < up_packed + 8 * B * m
    # a batch in which only some chunks qualify: QPs 10000.. get a second nonzero in one column
    b.CI[10000:, 2, 5] = 0.25
    ref = _oracle(restatement, b)
    r = gpu.solve_quadprog_batched_host(*b.fields(), want_active_set=True, chunk=4096)
    up_mixed, _ = gpu.last_transfer_bytes()
    assert_bit_identical(r._asdict(), ref, "mixed")
    assert up_packed < up_mixed < up_dense


def _retry_policy_expected(restatement, b, ref):
    """solveNextStep's TRY_TO_CONVERGE policy (src/mgqp.cpp:717-735) restated with the oracle: the second
    call runs without inequalities on the G object the first call overwrote with its Cholesky factor."""
    exp = {k: v.copy() for k, v in ref.items()}
    f = ref["f"]
    again = np.nonzero(((f != f) | np.isposinf(f)) & (ref["status"] != 2) & (ref["status"] != 3))[0]
    n, p, m = b.n, b.p, b.m
    for q in again:
        one = restatement.solve_one(b.G[q], b.g0[q], b.CE[q], b.ce0[q], b.CI[q], b.ci0[q], max_iter=default_max_iter(p, m))
        assert one["status"] == ref["status"][q]
        two = restatement.solve_one(one["G_factor"], b.g0[q], b.CE[q], b.ce0[q], np.zeros((n, 0)), np.zeros(0),
                                    max_iter=default_max_iter(p, 0))
        st, f2 = two["status"], two["f"]
        lost = st not in (2, 3) and (f2 != f2 or f2 == np.inf)
        exp["x"][q] = 0.0 if lost else two["x"]
        exp["f"][q] = f2
        exp["status"][q] = (1 if lost else st) | 16
        exp["iq"][q] = two["iq"]
        exp["iters"][q] = two["iters"]
        exp["A"][q] = 0x7fffffff
        exp["A"][q, :two["iq"]] = two["A"][:two["iq"]]
    return exp, again


@pytest.mark.parametrize("name,B,flags", [("C5", 6007, 0), ("C5", 6007, 8), ("S2", 3001, 0), ("C3", 2001, 0),
                                          ("R40", 257, 0)])
def test_infeasible_retry_policy_of_the_reference_adapter(gpu, restatement, name, B, flags):
    if name == "R40":                     # n = 40: CTA-per-QP kernel; the gather's Cholesky spans two lane strips
        b = workloads.make_random(B, 40, 2, 10, seed=40)
        b.CI[::3, :, 0] = 0.0; b.CI[::3, 0, 0] = 1.0; b.ci0[::3, 0] = -1.0     #  x0 >= 1
        b.CI[::3, :, 1] = 0.0; b.CI[::3, 0, 1] = -1.0; b.ci0[::3, 1] = -1.0    # -x0 >= 1
    else:
        b = workloads.make(name, B)
    if name == "S2":                      # the shipped shape is always feasible: contradict a bound pair
        b.ci0[::9, 3] = -7.0
        b.ci0[::9, 3 + 14] = -7.0
    if name == "C3":                      # general G, n = 14, p = 7: contradict the first joint's bound pair
        b.CI[::5, :, 0] = 0.0; b.CI[::5, 0, 0] = 1.0; b.ci0[::5, 0] = -3.0
        b.CI[::5, :, 1] = 0.0; b.CI[::5, 0, 1] = -1.0; b.ci0[::5, 1] = -3.0
    b.ce0[5::97, 0] = np.nan              # NaN cost both times: the command is zeroed (status 17)
    ref = _oracle(restatement, b)
    exp, again = _retry_policy_expected(restatement, b, ref)
    assert len(again) > 50 and (exp["status"] == 17).sum() >= B // 97 and (exp["status"] == 16).sum() > 20
    assert (exp["x"][exp["status"] == 17] == 0).all()
    got = _solve(gpu, b, flags=flags | gpu.FLAG_RETRY_WITHOUT_CI)
    assert_bit_identical(got, exp, name)
    assert ((got["status"] & 16) != 0).sum() == len(again)
    # the host-buffer entry point applies the same policy chunk by chunk
    r = gpu.solve_quadprog_batched_host(*b.fields(), want_active_set=True, chunk=1024, flags=flags | gpu.FLAG_RETRY_WITHOUT_CI)
    assert_bit_identical(r._asdict(), exp, name + " host")


def test_multi_device_entry_point(gpu, restatement):
    import torch
    devs = list(range(torch.cuda.device_count()))
    b = workloads.make("C5", 50001)
    r = gpu.solve_quadprog_batched_multi(*b.fields(), devices=devs, want_active_set=True, chunk=4096)
    assert_bit_identical(r._asdict(), _oracle(restatement, b), f"multi{len(devs)}")


def test_multi_device_entry_point_on_two_or_more_gpus(gpu, restatement):
    """solve_quadprog_batched_multi across real devices: needs a multi-GPU lease (skipped, not passed, without one).
    Every device must have taken chunks from the shared queue and the answers must be bit-identical."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    devs = list(range(torch.cuda.device_count()))
    b = workloads.make("C5", 400003)
    r = gpu.solve_quadprog_batched_multi(*b.fields(), devices=devs, want_active_set=True, chunk=8192)
    chunks = gpu.last_multi_chunks()
    assert len(chunks) == len(devs) and all(c > 0 for c in chunks) and sum(chunks) == -(-400003 // 8192)
    ref = _oracle(restatement, b)
    assert_bit_identical(r._asdict(), ref, f"multi{len(devs)}")


def test_single_qp_reference_semantics(gpu, restatement):
    """solve_quadprog(G, g0, CE, ce0, CI, ci0, x): cost returned, G overwritten by its factor,
    +inf when infeasible, the reference's two exception types."""
    for path in sorted(glob.glob(os.path.join(GOLDEN, "kat_K[1-6]*.npz"))):
        z = np.load(path)
        G = z["G"].copy()
        x = np.zeros(len(z["g0"]))
        st = int(z["status"])
        if st == 2:
            with pytest.raises(gpu.LogicError, match="Error in cholesky decomposition, sum: -1"):
                gpu.solve_quadprog(G, z["g0"], z["CE"], z["ce0"], z["CI"], z["ci0"], x)
        elif st == 3:
            with pytest.raises(RuntimeError, match="Constraints are linearly dependent"):
                gpu.solve_quadprog(G, z["g0"], z["CE"], z["ce0"], z["CI"], z["ci0"], x)
        else:
            f = gpu.solve_quadprog(G, z["g0"], z["CE"], z["ce0"], z["CI"], z["ci0"], x)
            assert np.array_equal(x, z["x"])
            assert f == float(z["f"])
            if st == 0:
                assert np.array_equal(G, z["G_factor"])


def test_split_division_matches_the_compilers_division(gpu):
    """The lane kernel's add_constraint sweep divides on the branch-free inline path of div.rn.f64
    (csrc/warp_prims.cuh); wherever that path accepts a pair its quotient must be the compiler's, bit for bit."""
    import ctypes as C
    from mgqp_b200 import _api
    L = _api.lib()
    L.mgqp_debug_div_check.restype = C.c_int
    L.mgqp_debug_div_check.argtypes = [C.c_int64, C.c_uint64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    for seed in (1, 0xC0FFEE):
        acc, bad = C.c_int64(-1), C.c_int64(-1)
        assert L.mgqp_debug_div_check(300_000_000, seed, C.byref(acc), C.byref(bad)) == 0
        assert bad.value == 0
        assert acc.value > 300_000_000      # two thirds of the draws are in the accepted range, two quotients each


def test_sparse_direction_takes_the_sweeps_fallback(gpu, restatement):
    """G = I with bound constraints: d = J' n has exact zeros, which the split division hands back to the general
    routine; the sweep's one-by-one fallback must give the reference's bits."""
    rng = np.random.default_rng(77)
    B, n, m = 5003, 7, 14
    G = np.broadcast_to(np.eye(n), (B, n, n)).copy()
    g0 = rng.standard_normal((B, n)) * 3
    CE = np.zeros((B, n, 1)); CE[:, 2, 0] = 1.0
    ce0 = rng.standard_normal((B, 1)) * 0.1
    CI = np.zeros((B, n, m))
    CI[:, np.arange(n), np.arange(n)] = 1.0
    CI[:, np.arange(n), n + np.arange(n)] = -1.0
    ci0 = np.abs(rng.standard_normal((B, m))) * 0.7 + 0.2
    b = workloads.QPBatch("sparse7", n, 1, m, G, g0, CE, ce0, CI, ci0)
    assert_bit_identical(_solve(gpu, b), _oracle(restatement, b), "sparse7")


def test_full_size_properties(gpu):
    """BASELINE config 2 at full size (1M QPs): size-independent properties instead of the oracle --
    primal feasibility, complementary active set, idempotence across a second launch."""
    import torch
    B = 1_000_000
    b = workloads.make("C2", B)
    t = _dev(b)
    r1 = gpu.solve_quadprog_batched(*t, want_active_set=True)
    r2 = gpu.solve_quadprog_batched(*t, want_active_set=True, flags=4)
    torch.cuda.synchronize()
    assert torch.equal(r1.x, r2.x) and torch.equal(r1.f, r2.f) and torch.equal(r1.A, r2.A)
    ok = r1.status == 0
    assert ok.float().mean().item() > 0.99
    x = r1.x
    s = torch.einsum("bnm,bn->bm", t[4], x) + t[5]
    e = torch.einsum("bnp,bn->bp", t[2], x) + t[3]
    assert s[ok].min().item() > -1e-9 and e[ok].abs().max().item() < 1e-9
    # every reported active inequality is tight
    A = r1.A[:, 1:8].long()
    valid = (A >= 0) & (A < b.m) & ok[:, None]
    tight = torch.gather(s, 1, A.clamp(0, b.m - 1)).abs()
    assert tight[valid].max().item() < 1e-9
