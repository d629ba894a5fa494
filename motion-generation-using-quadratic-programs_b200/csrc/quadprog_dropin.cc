// The reference's C++ entry point, re-implemented on top of the batched engine.
//
//   double solve_quadprog(Matrix<double>& G, Vector<double>& g0, const Matrix<double>& CE,
//                         const Vector<double>& ce0, const Matrix<double>& CI,
//                         const Vector<double>& ci0, Vector<double>& x);
//
// Same global-namespace symbol as lib/QuadProgpp/libquadprog.a (QuadProg++.hh:69-72), so
// src/mgqp.cpp:708/:725 links against libmgqp_b200.so unchanged.  A single QP is latency bound
// (two PCIe copies and a launch); this path exists for compatibility, the batch path is the product.
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
