// The reference adapter's infeasible-retry policy as a batch operation (MGQP_FLAG_RETRY_WITHOUT_CI).
//
// src/mgqp.cpp:717-735 (solveNextStep, TRY_TO_CONVERGE = true, :20): when solve_quadprog returns NaN
// or +inf the caller drops every inequality (CI.resize(0, n); ci0.resize(0)) and calls
// solve_quadprog again ON THE SAME G OBJECT -- which the first call has overwritten with its
// Cholesky factor, lower triangle mirrored into the upper one (QuadProg++.hh:42-45).  If that
// second answer is NaN / +inf as well the command is zeroed.
//
// Three small kernels around a second, ordinary launch of the solver on the compacted subset:
//   mark    -- list the QPs whose first answer is NaN / +inf;
//   gather  -- one warp per listed QP: G' = mirrored Cholesky factor of G in the reference's
//              operation order (cholesky_decomposition), plus copies of g0, CE, ce0;
//   scatter -- put the second answer (or the zero command) back in place, status |= RETRIED.
#pragma once
#include "warp_prims.cuh"

namespace mgqp {

__global__ void retry_mark_kernel(const double* __restrict__ f, const int* __restrict__ status, long long batch,
                                  long long* __restrict__ idx, unsigned long long* __restrict__ count) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= batch) return;
  const double v = f[q];
  const int st = status[q];
  // statuses 2 and 3 are the reference's exceptions: it never gets to the retry there
  const bool again = (st != 2 && st != 3) && (v != v || v == d_inf());
  if (again) idx[atomicAdd(count, 1ull)] = q;
}

struct RetryGather {
  int n, p;
  const double *G, *g0, *CE, *ce0;
  long long sG, sg0, sCE, sce0;
  double *G2, *g02, *CE2, *ce02;  // compact [count][...]
  const long long* idx;
  const unsigned long long* count;
};

__global__ void retry_gather_kernel(const __grid_constant__ RetryGather P) {
  const int lane = threadIdx.x & 31;
  const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  const int n = P.n, p = P.p;
  const long long cnt = (long long)*P.count;
  for (long long r = w; r < cnt; r += nw) {
    const long long q = P.idx[r];
    const double* G = P.G + q * P.sG;
    double* L = P.G2 + r * (long long)n * n;
    for (int e = lane; e < n * n; e += 32) L[e] = G[e];
    for (int e = lane; e < n; e += 32) P.g02[r * n + e] = P.g0[q * P.sg0 + e];
    for (int e = lane; e < n * p; e += 32) P.CE2[r * (long long)n * p + e] = P.CE[q * P.sCE + e];
    for (int e = lane; e < p; e += 32) P.ce02[r * p + e] = P.ce0[q * P.sce0 + e];
    __syncwarp();
    // cholesky_decomposition: row i, columns j >= i, k descending; lanes split j, never k.  The first
    // solve got through this factorisation (its status is not NOT_PD), so every pivot is positive or NaN.
    for (int i = 0; i < n; i++) {
      double mine[32];  // sums of this lane's columns j = i + lane + 32*t  (n <= 1024 per lane-strip)
      int nt = 0;
      for (int j = i + lane; j < n; j += 32) {
        double sum = L[i * n + j];
        for (int k = i - 1; k >= 0; k--) sum = d_sub(sum, d_mul(L[i * n + k], L[j * n + k]));
        if (nt < 32) mine[nt] = sum;
        nt++;
      }
      __syncwarp();
      if (lane == 0) L[i * n + i] = d_sqrt(mine[0]);
      __syncwarp();
      const double dii = L[i * n + i];
      nt = 0;
      for (int j = i + lane; j < n; j += 32) {
        if (j > i && nt < 32) L[j * n + i] = d_div(mine[nt], dii);
        nt++;
      }
      __syncwarp();
    }
    for (int e = lane; e < n * n; e += 32) {  // mirror: G[i][k] = G[k][i] for k > i
      const int i = e / n, k = e - i * n;
      if (k > i) L[e] = L[k * n + i];
    }
    __syncwarp();
  }
}

struct RetryScatter {
  int n, pm;
  const double *x2, *f2;
  const int *st2, *iq2, *A2, *it2;
  double *x, *f;
  int *status, *iq, *A, *iters;  // last three may be null
  const long long* idx;
  const unsigned long long* count;
  int p;
};

__global__ void retry_scatter_kernel(const __grid_constant__ RetryScatter P) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= (long long)*P.count) return;
  const long long q = P.idx[r];
  const double f2 = P.f2[r];
  int st = P.st2[r];
  const bool lost = (st != 2 && st != 3) && (f2 != f2 || f2 == d_inf());
  for (int k = 0; k < P.n; k++) P.x[q * P.n + k] = lost ? 0.0 : P.x2[r * P.n + k];  // *res = Zero(pbDOF), :732
  if (lost) st = 1;
  P.f[q] = f2;
  P.status[q] = st | 16;  // MGQP_STATUS_RETRIED
  if (P.iq) P.iq[q] = P.iq2[r];
  if (P.iters) P.iters[q] = P.it2[r];
  if (P.A) {
    for (int k = 0; k < P.pm; k++) P.A[q * P.pm + k] = (k < P.p) ? P.A2[r * P.p + k] : 0x7fffffff;
  }
}

}  // namespace mgqp
