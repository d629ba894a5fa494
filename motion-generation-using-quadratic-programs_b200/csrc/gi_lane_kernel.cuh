// Persistent kernel around gi_lane_solve: every WARP is an independent worker that pulls tiles of
// 32 consecutive QPs from the atomic work queue; lane q of the warp owns QP q of the tile.
//
// There is no staging pass: every lane reads its own QP's G, g0, CI and ci0 straight from global
// memory inside gi_lane_solve (consecutive 8-byte loads of one lane share 32-byte sectors, so HBM
// sees every byte once and L1 serves the rest); only x goes through shared memory on the way out
// so that its stores are coalesced.  (An L2 prefetch of the next tile was measured and removed: it
// doubled the DRAM read traffic -- the lines were evicted before use -- without changing the run
// time; profiles/r01_notes.md.)
#pragma once
#include "gi_lane.cuh"

namespace mgqp {

struct LaneParams {
  LaneView v;
  long long batch, tiles;
  const double* in[6];   // G, g0, CE, ce0, CI, ci0
  long long stride[6];
  const double* ci_val;  // column-compressed CI (mgqp_batch::CI_val / CI_row) or null
  const unsigned char* ci_row;
  long long s_ci_val, s_ci_row;
  double* x; double* f; int* status; int* iq; int* A; int* iters;
  int max_iter;
  unsigned* ticket;
  const unsigned long long* batch_dev;  // if set: the number of QPs lives on the device (<= batch)
};

template <int N, bool FMA>
__global__ void __launch_bounds__(256, 1) gi_lane_kernel(const __grid_constant__ LaneParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps_per_cta = blockDim.x >> 5;
  const LaneView& v = P.v;
  const int n = N, p = v.p, m = v.m, stride = v.stride;
  double* ws = reinterpret_cast<double*>(smem_raw) + (size_t)warp * 32 * stride;
  double* S = ws + lane * stride;
  const long long total_warps = (long long)gridDim.x * warps_per_cta;
  const long long my_warp = (long long)blockIdx.x * warps_per_cta + warp;

  auto next_tile = [&](long long prev) -> long long {
    if (P.ticket) {
      unsigned t = 0;
      if (lane == 0) t = atomicAdd(P.ticket, 1u);
      return (long long)__shfl_sync(MGQP_FULL, t, 0);
    }
    return (prev < 0) ? my_warp : prev + total_warps;
  };

  long long batch = P.batch, tiles = P.tiles;
  if (P.batch_dev) {
    const long long b = (long long)*P.batch_dev;
    batch = b < batch ? b : batch;
    tiles = (batch + 31) / 32;
  }
  long long tile = next_tile(-1);
  while (tile < tiles) {
    const long long q0 = tile * 32;
    const long long rem = batch - q0;
    const int cnt = rem < 32 ? (int)rem : 32;
    const long long tile_next = next_tile(tile);
    const long long myq = q0 + (lane < cnt ? lane : 0);
    LaneIO io;
    io.G = P.in[0] + myq * P.stride[0];
    io.g0 = P.in[1] + myq * P.stride[1];
    io.CE = P.in[2] + myq * P.stride[2];
    io.ce0 = P.in[3] + myq * P.stride[3];
    io.CI = P.in[4] + myq * P.stride[4];
    io.ci0 = P.in[5] + myq * P.stride[5];
    if (P.ci_val) {
      io.CIv = P.ci_val + myq * P.s_ci_val;
      io.CIr = P.ci_row + myq * P.s_ci_row;
    }
    double xo[N];
    double f;
    int status, iq, iters;
    gi_lane_solve<N, FMA>(v, S, io, P.max_iter, xo, f, status, iq, iters);
    __syncwarp();

    // ---- results: x through shared memory so that the global stores are coalesced
#pragma unroll
    for (int k = 0; k < N; k++) S[v.od + k] = xo[k];
    __syncwarp();
    for (int e = lane; e < cnt * n; e += 32) {
      const int q = e / n, k = e - q * n;
      P.x[q0 * n + e] = ws[q * stride + v.od + k];
    }
    if (lane < cnt) {
      P.f[q0 + lane] = f;
      P.status[q0 + lane] = status;
      if (P.iq) P.iq[q0 + lane] = iq;
      if (P.iters) P.iters[q0 + lane] = iters;
      if (P.A) {
        int* Ag = P.A + (q0 + lane) * (long long)(p + m);
        const signed char* Aq = reinterpret_cast<const signed char*>(S + v.oByte) + v.bA;
        for (int i = 0; i < p + m; i++) Ag[i] = (i < iq) ? (int)Aq[i] : 0x7fffffff;
      }
    }
    __syncwarp();
    tile = tile_next;
  }
}

}  // namespace mgqp
