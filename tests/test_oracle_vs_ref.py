"""CPU: pins the oracle (oracle/gi_oracle.c) to the reference.

1. against the committed golden fixtures, which were produced by the reference's shipped solver
   binary (tests/golden/make_golden.py) -- runs everywhere;
2. against that binary itself (oracle/_ref) on larger seeded batches -- runs wherever oracle/_ref
   has been built (the build container; the GPU box receives the prebuilt .so).
"""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from helpers import assert_bit_identical
from mgqp_b200 import workloads


def _load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "kat_K[1-6]*.npz"))))
def test_known_answers(restatement, path):
    z = np.load(path)
    r = restatement.solve_one(z["G"], z["g0"], z["CE"], z["ce0"], z["CI"], z["ci0"])
    assert r["status"] == int(z["status"])
    if r["status"] in (0, 1, 3):  # the binary leaves x as the last iterate in all three
        assert np.array_equal(r["x"], z["x"])
    if r["status"] == 0:
        assert r["f"] == float(z["f"])
        assert np.array_equal(r["G_factor"], z["G_factor"])  # "G is modified" (QuadProg++.hh:42-45)
    if r["status"] == 1:
        assert np.isposinf(r["f"])
    if r["status"] == 2:  # message carries the pivot: "Error in cholesky decomposition, sum: -1"
        assert str(z["message"]) == "Error in cholesky decomposition, sum: %g" % r["f"]


def test_known_answer_values():
    """The numbers SURVEY.md 8c records for K1..K3 (what the binary printed during the survey)."""
    z = _load("kat_K1_demo.npz")
    assert float(z["f"]) == 12.0 and np.allclose(z["x"], [1.0, 2.0], atol=1e-14)
    z = _load("kat_K2_unconstrained.npz")
    assert float(z["f"]) == -6.0 and np.allclose(z["x"], [-2.0, -1.0], atol=1e-14)
    z = _load("kat_K3_least_norm.npz")
    assert float(z["f"]) == 4.5 and np.allclose(z["x"], [1.0, 2.0, 2.0], atol=1e-14)
    z = _load("kat_K7_dim_mismatch.npz")
    assert str(z["message"]) == "The matrix CE is incompatible (incorrect number of rows 3 , expecting 2)"


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "batch_[CSR]*[0-9].npz"))))
def test_golden_batches(restatement, path):
    z = np.load(path)
    B, n = z["g0"].shape
    p, m = z["ce0"].shape[1], z["ci0"].shape[1]
    r = restatement.solve_batch(n, p, m, z["G"], z["g0"], z["CE"], z["ce0"], z["CI"], z["ci0"], threads=2)
    assert_bit_identical(r, dict(x=z["x"], f=z["f"], status=z["status"]), os.path.basename(path))


def test_golden_mpc_shape(restatement):
    z = _load("batch_C4_seedonly.npz")
    b = workloads.make("C4", 1)
    r = restatement.solve_batch(b.n, b.p, b.m, *b.fields())
    assert_bit_identical(r, dict(x=z["x"], f=z["f"], status=z["status"]), "C4")


@pytest.mark.parametrize("name,B", [("C2", 4000), ("C5", 4000), ("C3", 1500), ("S0", 1500), ("S2", 1500)])
def test_against_shipped_binary(restatement, reference_binary, name, B):
    if reference_binary is None:
        pytest.skip("oracle/_ref not built here")
    b = workloads.make(name, B, seed=97 + B)
    ref = reference_binary.solve_batch(b.n, b.p, b.m, *b.fields(), threads=4)
    got = restatement.solve_batch(b.n, b.p, b.m, *b.fields(), threads=4)
    assert_bit_identical(got, ref, name)


@pytest.mark.parametrize("n,p,m", [(1, 0, 2), (2, 2, 0), (6, 3, 30), (11, 0, 25), (24, 6, 60), (48, 5, 120)])
def test_random_shapes_against_shipped_binary(restatement, reference_binary, n, p, m):
    if reference_binary is None:
        pytest.skip("oracle/_ref not built here")
    b = workloads.make_random(200, n, p, m, seed=n + 13 * m)
    ref = reference_binary.solve_batch(b.n, b.p, b.m, *b.fields(), threads=4)
    got = restatement.solve_batch(b.n, b.p, b.m, *b.fields(), threads=4)
    assert_bit_identical(got, ref, f"R{n}_{p}_{m}")


def test_exception_cases_against_shipped_binary(restatement, reference_binary):
    if reference_binary is None:
        pytest.skip("oracle/_ref not built here")
    rng = np.random.default_rng(5)
    for trial in range(50):
        n = int(rng.integers(2, 9))
        b = workloads.make_random(1, n, 2, 4, seed=trial)
        G, g0, CE, ce0, CI, ci0 = (a[0].copy() for a in b.fields())
        kind = trial % 3
        if kind == 0:
            CE[:, 1] = CE[:, 0] * 2.0  # dependent equalities -> runtime_error
        elif kind == 1:
            G = G - 3.0 * np.eye(n)  # indefinite -> logic_error
        ref = reference_binary.solve_one(G, g0, CE, ce0, CI, ci0)
        got = restatement.solve_one(G, g0, CE, ce0, CI, ci0)
        assert got["status"] == ref["status"], (trial, kind)
        if ref["status"] == 2:
            assert ref["message"] == "Error in cholesky decomposition, sum: %g" % got["f"]
        else:
            assert np.array_equal(got["x"], ref["x"])


def test_retry_sequence_of_the_adapter_against_the_binary(restatement, reference_binary):
    """solveNextStep (src/mgqp.cpp:708-735) calls the solver a second time on the SAME G object, which the
    first call has overwritten with its mirrored Cholesky factor, after dropping the inequalities.  The
    oracle's restatement of that two-call sequence (what MGQP_FLAG_RETRY_WITHOUT_CI is tested against on
    the GPU) must match the shipped binary run the same way, bit for bit."""
    if reference_binary is None:
        pytest.skip("oracle/_ref not built here")
    b = workloads.make("C5", 1500)           # general G (not the identity): the re-factorisation matters
    n, p, m = b.n, b.p, b.m
    tried = 0
    for q in range(b.batch):
        one = reference_binary.solve_one(b.G[q], b.g0[q], b.CE[q], b.ce0[q], b.CI[q], b.ci0[q])
        if not (np.isposinf(one["f"]) or np.isnan(one["f"])):
            continue
        tried += 1
        mine = restatement.solve_one(b.G[q], b.g0[q], b.CE[q], b.ce0[q], b.CI[q], b.ci0[q])
        assert np.array_equal(mine["G_factor"], one["G_factor"])     # what the second call gets as G
        assert not np.array_equal(one["G_factor"], b.G[q])
        two = reference_binary.solve_one(one["G_factor"], b.g0[q], b.CE[q], b.ce0[q], np.zeros((n, 0)), np.zeros(0))
        mine2 = restatement.solve_one(mine["G_factor"], b.g0[q], b.CE[q], b.ce0[q], np.zeros((n, 0)), np.zeros(0))
        assert two["status"] == mine2["status"]
        assert np.array_equal(two["x"], mine2["x"]) and two["f"] == mine2["f"]
    assert tried > 100
