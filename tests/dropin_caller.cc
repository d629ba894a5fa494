// TEST INFRASTRUCTURE: a translation unit written exactly like the reference's caller
// (src/mgqp.cpp:695-708, :723-725: resize / operator[] fill / t(CE) / solve_quadprog).  Compiled against this
// repository's include/QuadProgpp headers by tests/test_dropin_cpp.py and -- where /root/reference exists --
// against the REFERENCE's own include/QuadProgpp/{QuadProg++.hh,Array.hh} by oracle/Makefile
// (oracle/_ref/dropin_caller_refhdr), both linked to libmgqp_b200.so.
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
