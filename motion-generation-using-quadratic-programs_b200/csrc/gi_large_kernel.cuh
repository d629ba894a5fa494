// Persistent CTA-per-QP kernel around gi_large_solve: each CTA pulls QP indices from the atomic
// work queue and owns one slice of the global workspace (Cholesky factor | J | R | s).
#pragma once
#include "gi_large.cuh"

namespace mgqp {

struct LargeParams {
  LargeView v;
  long long batch;
  const double* in[6];
  long long stride[6];
  double* x; double* f; int* status; int* iq; int* A; int* iters;
  int max_iter;
  unsigned* ticket;
  double* workspace;  // gridDim.x * v.ws_doubles
  unsigned* gate;     // admission counter of the set-up phase (zeroed before the launch) or null
  int gate_limit;
  const unsigned long long* batch_dev;  // if set: the number of QPs lives on the device (<= batch)
};

template <bool FMA>
// CTA shape.  Half of this kernel's stall cycles are block barriers: threads waiting for thread 0's serial chains (the
// hypotenuse chain of add_constraint, the back-substitution of update_r -- both fixed by the reference's operation
// order) or for the slowest row of a pass.  What fills those waits is OTHER QPs on the same SM, so the CTA is made
// small: 128 threads, 96 registers (five CTAs per SM by registers), 49 KB of shared memory for the 350-variable shape
// (MGQP_ROWBLK = 1), which admits FOUR of them.  Measured on that shape (2368 QPs, exact / MGQP_FLAG_FAST_MATH QP/s;
// profiles/r02_notes.md):
//   256 x 2: 4 221 / 4 683     384 x 2: 4 434 / 5 042     256 x 3: 4 615 / 5 130     224 x 3: 4 740 / 5 297
//   192 x 4: 5 343 / 5 942     160 x 4: 4 532 / 5 049     128 x 4: 5 449 / 6 080
// (With x_old / u_old moved to the global workspace the CTA shrinks to 43 KB and five fit: 5 156 / 5 748 against
// 5 107 / 5 661 for four on the same box -- within the noise, not kept.)
#ifndef MGQP_LARGE_BOUNDS_T
#define MGQP_LARGE_BOUNDS_T 128
#define MGQP_LARGE_BOUNDS_B 5
#endif
__global__ void __launch_bounds__(MGQP_LARGE_BOUNDS_T, MGQP_LARGE_BOUNDS_B) gi_large_kernel(const __grid_constant__ LargeParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ long long s_next;
  double* ws = P.workspace + (size_t)blockIdx.x * P.v.ws_doubles;
  long long q = -1;
  long long batch = P.batch;
  if (P.batch_dev) {
    const long long b = (long long)*P.batch_dev;
    batch = b < batch ? b : batch;
  }
  for (;;) {
    if (threadIdx.x == 0) {
      if (P.ticket) s_next = (long long)atomicAdd(P.ticket, 1u);
      else s_next = (q < 0) ? (long long)blockIdx.x : q + gridDim.x;
    }
    __syncthreads();
    q = s_next;
    __syncthreads();
    if (q >= batch) break;
    LaneIO io;
    io.G = P.in[0] + q * P.stride[0];
    io.g0 = P.in[1] + q * P.stride[1];
    io.CE = P.in[2] + q * P.stride[2];
    io.ce0 = P.in[3] + q * P.stride[3];
    io.CI = P.in[4] + q * P.stride[4];
    io.ci0 = P.in[5] + q * P.stride[5];
    LargeOut out;
    out.x = P.x + q * P.v.n;
    out.f = P.f + q;
    out.status = P.status + q;
    out.iq = P.iq ? P.iq + q : nullptr;
    out.A = P.A ? P.A + q * (long long)(P.v.p + P.v.m) : nullptr;
    out.iters = P.iters ? P.iters + q : nullptr;
    out.gate = P.gate;
    out.gate_limit = P.gate_limit;
    gi_large_solve<FMA>(P.v, smem_raw, ws, io, P.max_iter, out);
  }
}

}  // namespace mgqp
