"""The C++ drop-in: a translation unit written exactly like the reference's caller
(src/mgqp.cpp:695-708: resize / operator[] fill / t(CE) / solve_quadprog) must compile against
include/QuadProgpp/QuadProg++.hh and link against libmgqp_b200.so.  Compile + link run on CPU;
executing it needs a GPU."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

SRC = r'''
#include <cstdio>
#include <stdexcept>
#include <cmath>
#include "QuadProg++.hh"
int main() {
  // K1 of SURVEY.md 8c, assembled the way solveNextStep does: rows = constraints, then t()
  ArrayHH::Matrix<double> G, CE, CI; ArrayHH::Vector<double> g0, ce0, ci0, x;
  G.resize(2, 2); G[0][0] = 4; G[0][1] = -2; G[1][0] = -2; G[1][1] = 4;
  g0.resize(2); g0[0] = 6; g0[1] = 0;
  CE.resize(1, 2); CE[0][0] = 1; CE[0][1] = 1; ce0.resize(1); ce0[0] = -3;
  CI.resize(3, 2); CI[0][0] = 1; CI[0][1] = 0; CI[1][0] = 0; CI[1][1] = 1; CI[2][0] = 1; CI[2][1] = 1;
  ci0.resize(3); ci0[0] = 0; ci0[1] = 0; ci0[2] = -2;
  x.resize(2);
  double f = solve_quadprog(G, g0, ArrayHH::t(CE), ce0, ArrayHH::t(CI), ci0, x);
  std::printf("f=%.17g x=%.17g %.17g G00=%.17g\n", f, x[0], x[1], G[0][0]);
  // retry path of src/mgqp.cpp:723-725: inequalities dropped
  CI.resize(0, 2); ci0.resize(0);
  ArrayHH::Matrix<double> G2; G2.resize(2, 2); G2[0][0] = 4; G2[0][1] = -2; G2[1][0] = -2; G2[1][1] = 4;
  f = solve_quadprog(G2, g0, ArrayHH::t(CE), ce0, ArrayHH::t(CI), ci0, x);
  std::printf("f2=%.17g\n", f);
  try { ArrayHH::Matrix<double> B; B.resize(3, 1); solve_quadprog(G2, g0, B, ce0, ArrayHH::t(CI), ci0, x); }
  catch (const std::logic_error& e) { std::printf("logic_error: %s\n", e.what()); }
  return 0;
}
'''


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    import __graft_entry__ as g
    g.build()
    d = tmp_path_factory.mktemp("dropin")
    src = d / "caller.cc"
    src.write_text(SRC)
    out = d / "caller"
    pkg = os.path.join(ROOT, "motion-generation-using-quadratic-programs_b200")
    subprocess.check_call(["g++", "-std=gnu++14", "-O1", "-I", os.path.join(ROOT, "include", "QuadProgpp"), str(src),
                           "-L", pkg, "-lmgqp_b200", f"-Wl,-rpath,{pkg}", "-o", str(out)])
    return str(out)


def test_reference_style_caller_compiles_and_links(exe):
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_reference_style_caller_runs(exe):
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    z = np.load(os.path.join(ROOT, "tests", "golden", "kat_K1_demo.npz"))
    line = out.splitlines()[0]
    f = float(line.split("f=")[1].split()[0])
    xs = [float(v) for v in line.split("x=")[1].split()[:2]]
    assert f == float(z["f"]) and xs == list(z["x"])
    assert "G00=2" in line                      # G overwritten by its Cholesky factor
    assert "logic_error: The matrix CE is incompatible (incorrect number of rows 3 , expecting 2)" in out
