// Set-up of the CTA-per-QP solver on the fp64 tensor cores (MGQP_FLAG_FAST_MATH, n >= 64): blocked Cholesky
// G = L L' and blocked triangular inversion X = L^-1, tile size 8, every 8x8x8 tile product two
// mma.sync.aligned.m8n8k4.f64 (SASS: DMMA).
//
// Replaces, for the horizon-stacked shapes (BASELINE.json configs[3]: n = 350), the set-up section of the
// reference's solve_quadprog -- cholesky_decomposition, the n forward_eliminations that build J = L^-T, and
// cholesky_solve for x0 = -G^-1 g0 (QuadProg++.o +0x2df0, +0x837, +0x96a) -- which in the
// reference's summation order is a chain of n^2/2 dependent multiply-adds per row of J and took three quarters of
// this kernel's time (profiles/r02_notes.md).  The blocked forms change the order in which each sum is accumulated,
// so this path exists only behind MGQP_FLAG_FAST_MATH: verdicts and active sets as the reference, x and f within
// 1e-9 relative (BASELINE.json north_star), not bit-identical.
//
// The solver stores J transposed (element [r][i] = J[i][r]), and J' = (L^-T)' = L^-1: X IS that array, row-major,
// lower triangular.  With X at hand x0 needs no triangular solve: y = X g0, x0 = -X' y.
#pragma once
#include <cuda_runtime.h>

namespace mgqp {

// D(8x8) = A(8x4) B(4x8) + C.  Fragments (PTX ISA, m8n8k4 .f64): lane = 4*g + t;  a = A[g][t], b = B[t][g],
// c0 = C[g][2t], c1 = C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// A(n x n, leading dimension n) read with an identity border up to the next multiple of 8
__device__ __forceinline__ double dm_get(const double* A, int n, int r, int c) {
  return (r < n && c < n) ? A[(size_t)r * n + c] : ((r == c) ? 1.0 : 0.0);
}

// Returns false (and the offending pivot) if G is not positive definite.  On success: Lw = L (lower triangle valid),
// Xw = L^-1 (full array, zeros above the diagonal).  sbuf: >= 8 * (n + 16) + 64 * (warps + 2) doubles of shared memory.
__device__ inline bool large_setup_dmma(int n, const double* __restrict__ G, double* __restrict__ Lw,
                                        double* __restrict__ Xw, double* __restrict__ sbuf, double* bad_pivot) {
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, W = T >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int nb = (n + 7) >> 3;
  double* sD = sbuf;                 // 64: diagonal tile / its inverse
  double* sP = sbuf + 64;            // panel: (8 nb) x 8
  double* sT = sP + (size_t)8 * (nb + 1) * 8;   // per-warp 64-double transposition tiles
  __shared__ int s_fail;
  // G's upper triangle (the one the reference reads) mirrored into a full copy; X = 0
  for (int e = tid; e < n * n; e += T) {
    const int r = e / n, c = e - r * n;
    Lw[e] = (c >= r) ? G[e] : G[(size_t)c * n + r];
    Xw[e] = 0.0;
  }
  if (tid == 0) s_fail = 0;
  __syncthreads();
  // ------------------------------------------------------------ blocked right-looking Cholesky
  for (int kb = 0; kb < nb; kb++) {
    const int k0 = 8 * kb;
    if (tid < 64) sD[tid] = dm_get(Lw, n, k0 + (tid >> 3), k0 + (tid & 7));
    __syncthreads();
    if (tid == 0) {  // 8x8 Cholesky of the diagonal tile, in place (lower)
      for (int i = 0; i < 8 && !s_fail; i++) {
        for (int j = 0; j <= i; j++) {
          double sum = sD[i * 8 + j];
          for (int k = 0; k < j; k++) sum = fma(-sD[i * 8 + k], sD[j * 8 + k], sum);
          if (i == j) {
            if (!(sum > 0.0) && !(sum != sum)) {  // the reference fails on sum <= 0; a NaN pivot passes its test
              s_fail = 1;
              *bad_pivot = sum;
              break;
            }
            sD[i * 8 + i] = sqrt(sum);
          } else {
            sD[i * 8 + j] = sum / sD[j * 8 + j];
          }
        }
      }
    }
    __syncthreads();
    if (s_fail) return false;
    if (tid < 64) {
      const int r = k0 + (tid >> 3), c = k0 + (tid & 7);
      if (r < n && c < n && c <= r) Lw[(size_t)r * n + c] = sD[tid];
    }
    // panel: L[r][k0..k0+7] = A[r][k0..] L_kk^-T, one row per thread; kept in shared memory for the update
    for (int r = k0 + 8 + tid; r < 8 * nb; r += T) {
      double v[8];
#pragma unroll
      for (int c = 0; c < 8; c++) v[c] = dm_get(Lw, n, r, k0 + c);
#pragma unroll
      for (int c = 0; c < 8; c++) {
        double a = v[c];
#pragma unroll
        for (int q = 0; q < 8; q++)
          if (q < c) a = fma(-v[q], sD[c * 8 + q], a);
        v[c] = a / sD[c * 8 + c];
      }
#pragma unroll
      for (int c = 0; c < 8; c++) {
        sP[(size_t)(r - k0 - 8) * 8 + c] = v[c];
        if (r < n && k0 + c < n) Lw[(size_t)r * n + k0 + c] = v[c];
      }
    }
    __syncthreads();
    // trailing update: A_ij -= L_ik L_jk' for the tiles bj <= bi behind the panel
    const int rem = nb - kb - 1;
    const int ntile = rem * (rem + 1) / 2;
    // (four tiles per trip: their eight C loads are in flight together -- the matrices live in L2 / DRAM)
    for (int tl0 = warp; tl0 < ntile; tl0 += 4 * W) {
      int rr[4], cc[4], pi[4], pj[4];
      double c0[4], c1[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int tl = tl0 + u * W;
        // (bi, bj) from the linear index of the lower triangle
        int bi = (int)((sqrtf(8.0f * (float)tl + 1.0f) - 1.0f) * 0.5f);
        while ((bi + 1) * (bi + 2) / 2 <= tl) bi++;
        while (bi * (bi + 1) / 2 > tl) bi--;
        const int bj = tl - bi * (bi + 1) / 2;
        pi[u] = 8 * bi; pj[u] = 8 * bj;
        rr[u] = (tl < ntile) ? k0 + 8 + 8 * bi + g : n;   // r = n: nothing to load or store
        cc[u] = k0 + 8 + 8 * bj + 2 * t;
        c0[u] = dm_get(Lw, n, rr[u], cc[u]);
        c1[u] = dm_get(Lw, n, rr[u], cc[u] + 1);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (tl0 + u * W < ntile) {  // warp-uniform
          const double* Pi = sP + (size_t)pi[u] * 8;
          const double* Pj = sP + (size_t)pj[u] * 8;
          dmma884(c0[u], c1[u], -Pi[g * 8 + t], Pj[g * 8 + t]);
          dmma884(c0[u], c1[u], -Pi[g * 8 + t + 4], Pj[g * 8 + t + 4]);
          if (rr[u] < n) {
            if (cc[u] < n) Lw[(size_t)rr[u] * n + cc[u]] = c0[u];
            if (cc[u] + 1 < n) Lw[(size_t)rr[u] * n + cc[u] + 1] = c1[u];
          }
        }
      }
    }
    __syncthreads();
  }
  // ------------------------------------------------------------ X = L^-1, block row by block row
  for (int bi = 0; bi < nb; bi++) {
    const int i0 = 8 * bi;
    if (tid < 64) {
      const int r = i0 + (tid >> 3), c = i0 + (tid & 7);
      sD[tid] = (c <= r) ? dm_get(Lw, n, r, c) : 0.0;
    }
    __syncthreads();
    if (tid < 8) {  // column tid of L_ii^-1 by forward substitution, into sD + 64
      double x[8];
      for (int i = 0; i < 8; i++) {
        double a = (i == tid) ? 1.0 : 0.0;
        for (int k = 0; k < i; k++) a = fma(-sD[i * 8 + k], x[k], a);
        x[i] = a / sD[i * 8 + i];
        sP[i * 8 + tid] = x[i];     // sP[0..63] doubles as the inverse tile in this phase
      }
    }
    __syncthreads();
    if (tid < 64) {
      const int r = i0 + (tid >> 3), c = i0 + (tid & 7);
      if (r < n && c < n) Xw[(size_t)r * n + c] = sP[tid];
    }
    double* myT = sT + (size_t)warp * 64;
    for (int bj = warp; bj < bi; bj += W) {
      const int j0 = 8 * bj;
      // S = sum_{bk = bj}^{bi-1} L(bi, bk) X(bk, bj)
      double s0 = 0.0, s1 = 0.0;
      for (int bk0 = bj; bk0 < bi; bk0 += 4) {  // four k-blocks per trip: sixteen operand loads in flight
        double a[4][2], b[4][2];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int k0 = 8 * (bk0 + u);
          const bool on = bk0 + u < bi;
          a[u][0] = on ? dm_get(Lw, n, i0 + g, k0 + t) : 0.0;
          a[u][1] = on ? dm_get(Lw, n, i0 + g, k0 + t + 4) : 0.0;
          b[u][0] = on ? dm_get(Xw, n, k0 + t, j0 + g) : 0.0;
          b[u][1] = on ? dm_get(Xw, n, k0 + t + 4, j0 + g) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (bk0 + u < bi) {  // warp-uniform
            dmma884(s0, s1, a[u][0], b[u][0]);
            dmma884(s0, s1, a[u][1], b[u][1]);
          }
        }
      }
      // X(bi, bj) = -L_ii^-1 S: S goes through shared memory to become a B operand
      myT[g * 8 + 2 * t] = s0;
      myT[g * 8 + 2 * t + 1] = s1;
      __syncwarp();
      double x0 = 0.0, x1 = 0.0;
      dmma884(x0, x1, -sP[g * 8 + t], myT[t * 8 + g]);
      dmma884(x0, x1, -sP[g * 8 + t + 4], myT[(t + 4) * 8 + g]);
      __syncwarp();
      const int r = i0 + g, c = j0 + 2 * t;
      if (r < n) {
        if (c < n) Xw[(size_t)r * n + c] = x0;
        if (c + 1 < n) Xw[(size_t)r * n + c + 1] = x1;
      }
    }
    __syncthreads();
  }
  return true;
}

}  // namespace mgqp
