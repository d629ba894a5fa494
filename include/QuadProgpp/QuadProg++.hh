// Drop-in for the reference's include/QuadProgpp/QuadProg++.hh (prototype at :69-72).
//
//   min 0.5 * x G x + g0 x     s.t.  CE^T x + ce0 = 0,   CI^T x + ci0 >= 0
//   G: n*n, g0: n, CE: n*p, ce0: p, CI: n*m, ci0: m, x: n          (QuadProg++.hh:8-25)
//
// Returns the cost, or +infinity if the problem is infeasible (x is then not meaningful).
// Throws std::logic_error for dimension mismatches and a non-positive-definite G, and
// std::runtime_error("Constraints are linearly dependent"), like the reference binary.
// G is overwritten by its Cholesky factor.  The work is done on the GPU by libmgqp_b200.so
// (include/mgqp_b200.h); for batches call solve_quadprog_batched() directly.
#ifndef MGQP_B200_QUADPROGPP_COMPAT
#define MGQP_B200_QUADPROGPP_COMPAT

#include "Array.hh"

using namespace ArrayHH;

double solve_quadprog(Matrix<double>& G, Vector<double>& g0, const Matrix<double>& CE,
                      const Vector<double>& ce0, const Matrix<double>& CI, const Vector<double>& ci0,
                      Vector<double>& x);

#endif
