"""CPU: the product library loads and exports every entry point include/mgqp_b200.h declares,
plus the reference's own C++ symbol for the drop-in path.  No compute call is made."""
import ctypes as C
import os
import re
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as g
    g.build()
    import mgqp_b200
    return mgqp_b200.lib_path()


def _declared():
    hdr = open(os.path.join(ROOT, "include", "mgqp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(solve_quadprog_batched\w*|mgqp_[a-z_]+)\s*\(", hdr)))


def test_exports_every_declared_symbol(built_lib):
    lib = C.CDLL(built_lib)
    names = _declared()
    assert "solve_quadprog_batched" in names and "mgqp_solve_one" in names and len(names) >= 9
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/mgqp_b200.h but not exported"


def test_exports_reference_cxx_symbol(built_lib):
    """src/mgqp.cpp:708 links against this exact mangled name."""
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True).stdout
    assert "_Z14solve_quadprogRN7ArrayHH6MatrixIdEERNS_6VectorIdEERKS1_RKS4_S7_S9_S5_" in out


def test_plan_and_bytes(built_lib):
    import mgqp_b200
    assert mgqp_b200.algorithmic_bytes(7, 1, 42) == 3268
    assert mgqp_b200.algorithmic_bytes(14, 7, 60) == 9844
    assert mgqp_b200.algorithmic_bytes(14, 10, 28) == 6364
    gw, warps, smem = mgqp_b200.plan(7, 1, 42)
    assert gw == 1 and warps >= 2 and smem <= 227 * 1024      # one QP per lane
    assert mgqp_b200.plan(14, 10, 28)[0] == 16
    assert mgqp_b200.plan(32, 4, 70)[0] == 32


def test_no_cpu_fallback_message(built_lib):
    """Without a GPU the compute entry points must fail loudly, not fall back."""
    import numpy as np
    import torch
    import mgqp_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    z = np.zeros
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        mgqp_b200.solve_quadprog_batched_host(np.eye(2)[None], z((1, 2)), z((1, 2, 0)), z((1, 0)), z((1, 2, 0)), z((1, 0)))


def test_dimension_errors_match_reference_text():
    import numpy as np
    import mgqp_b200
    with pytest.raises(mgqp_b200.LogicError, match=r"The matrix CE is incompatible \(incorrect number of rows 3 , expecting 2\)"):
        mgqp_b200.solve_quadprog(np.eye(2), np.zeros(2), np.ones((3, 1)), np.zeros(1), np.zeros((2, 0)), np.zeros(0))
    with pytest.raises(mgqp_b200.LogicError, match=r"The vector ci0 is incompatible"):
        mgqp_b200.solve_quadprog(np.eye(2), np.zeros(2), np.zeros((2, 0)), np.zeros(0), np.zeros((2, 3)), np.zeros(2))
