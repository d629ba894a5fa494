// The reference's C++ entry point, re-implemented on top of the batched engine.
//
//   double solve_quadprog(Matrix<double>& G, Vector<double>& g0, const Matrix<double>& CE,
//                         const Vector<double>& ce0, const Matrix<double>& CI,
//                         const Vector<double>& ci0, Vector<double>& x);
//
// Same global-namespace symbol as lib/QuadProgpp/libquadprog.a (QuadProg++.hh:69-72), so
// src/mgqp.cpp:708/:725 links against libmgqp_b200.so unchanged.  A single QP is latency bound
// (two PCIe copies and a launch); this path exists for compatibility, the batch path is the product.
#include <chrono>
#include <cstdlib>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <vector>

#include "../../include/QuadProgpp/QuadProg++.hh"
#include "../../include/mgqp_b200.h"

double solve_quadprog(Matrix<double>& G, Vector<double>& g0, const Matrix<double>& CE,
                      const Vector<double>& ce0, const Matrix<double>& CI, const Vector<double>& ci0,
                      Vector<double>& x) {
  std::ostringstream msg;
  const unsigned int n = G.ncols(), p = CE.ncols(), m = CI.ncols();
  // the five checks of the reference, with its exception texts (.rodata of QuadProg++.o)
  if (G.nrows() != n) {
    msg << "The matrix G is not a squared matrix (" << G.nrows() << " x " << G.ncols() << ")";
    throw std::logic_error(msg.str());
  }
  if (CE.nrows() != n) {
    msg << "The matrix CE is incompatible (incorrect number of rows " << CE.nrows() << " , expecting " << n << ")";
    throw std::logic_error(msg.str());
  }
  if (ce0.size() != p) {
    msg << "The vector ce0 is incompatible (incorrect dimension " << ce0.size() << ", expecting " << p << ")";
    throw std::logic_error(msg.str());
  }
  if (CI.nrows() != n) {
    msg << "The matrix CI is incompatible (incorrect number of rows " << CI.nrows() << " , expecting " << n << ")";
    throw std::logic_error(msg.str());
  }
  if (ci0.size() != m) {
    msg << "The vector ci0 is incompatible (incorrect dimension " << ci0.size() << ", expecting " << m << ")";
    throw std::logic_error(msg.str());
  }
  x.resize(n);
  if (n == 0) return 0.0;

  // Matrix storage is one contiguous row-major block behind v[0] (Array.hh:910-919): no packing.
  std::vector<double> factor((size_t)n * n), xs(n);
  double f = 0.0;
  int32_t status = 0;
  int device = 0;
  if (const char* e = std::getenv("MGQP_DEVICE")) device = std::atoi(e);
  const int rc = mgqp_solve_one((int)n, (int)p, (int)m, G[0], &g0[0], p ? CE[0] : nullptr, p ? &ce0[0] : nullptr,
                                m ? CI[0] : nullptr, m ? &ci0[0] : nullptr, xs.data(), &f, &status, factor.data(),
                                device);
  if (rc != MGQP_OK) throw std::runtime_error(std::string("mgqp_b200: ") + mgqp_last_error());
  if (status == MGQP_STATUS_NOT_PD) {
    msg << "Error in cholesky decomposition, sum: " << f;
    throw std::logic_error(msg.str());
  }
  for (unsigned int i = 0; i < n; i++)
    for (unsigned int j = 0; j < n; j++) G[i][j] = factor[(size_t)i * n + j];
  for (unsigned int i = 0; i < n; i++) x[i] = xs[i];
  if (status == MGQP_STATUS_EQ_DEPENDENT) throw std::runtime_error("Constraints are linearly dependent");
  if (status == MGQP_STATUS_INFEASIBLE) return std::numeric_limits<double>::infinity();
  return f;
}

// Bench hook: wall-clock microseconds of `reps` consecutive calls of the drop-in solve_quadprog() above on ONE
// QP given in QuadProg++ layout (G is rebuilt before every call, outside the timed region, because the call
// overwrites it with its Cholesky factor).  Returns 0, or -1 if a call threw.
extern "C" int mgqp_dropin_latency(int32_t n, int32_t p, int32_t m, const double* G, const double* g0, const double* CE,
                                   const double* ce0, const double* CI, const double* ci0, int reps, double* us_out,
                                   double* x_out, double* f_out) {
  try {
    Matrix<double> Gm(n, n), CEm(n, p), CIm(n, m);
    Vector<double> g0v(n), ce0v(p), ci0v(m), xv(n);
    for (int i = 0; i < n; i++) g0v[i] = g0[i];
    for (int i = 0; i < n; i++)
      for (int j = 0; j < p; j++) CEm[i][j] = CE[(size_t)i * p + j];
    for (int j = 0; j < p; j++) ce0v[j] = ce0[j];
    for (int i = 0; i < n; i++)
      for (int j = 0; j < m; j++) CIm[i][j] = CI[(size_t)i * m + j];
    for (int j = 0; j < m; j++) ci0v[j] = ci0[j];
    double f = 0.0;
    for (int r = 0; r < reps; r++) {
      for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) Gm[i][j] = G[(size_t)i * n + j];
      const auto t0 = std::chrono::steady_clock::now();
      f = solve_quadprog(Gm, g0v, CEm, ce0v, CIm, ci0v, xv);
      const auto t1 = std::chrono::steady_clock::now();
      us_out[r] = std::chrono::duration<double, std::micro>(t1 - t0).count();
    }
    if (x_out) for (int i = 0; i < n; i++) x_out[i] = xv[i];
    if (f_out) *f_out = f;
    return 0;
  } catch (...) {
    return -1;
  }
}
