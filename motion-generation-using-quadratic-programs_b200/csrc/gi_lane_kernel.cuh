// Persistent kernel around gi_lane_tile: every WARP is an independent worker that pulls tiles of
// 32 consecutive QPs from the atomic work queue; lane q of the warp owns QP q of the tile.
//
// The warp stages its tile cooperatively (gi_lane.cuh, phases A and B): consecutive lanes read consecutive
// doubles of the tile's G, g0, CI and ci0 blocks, so every global load is coalesced and HBM sees every byte
// once; x goes through shared memory on the way out so that its stores are coalesced, too.
#pragma once
#include "gi_lane.cuh"

#ifndef MGQP_LANE_CTA_SYNC
#define MGQP_LANE_CTA_SYNC 4
#endif
#ifndef MGQP_LANE_SYNC_ROUNDS
#define MGQP_LANE_SYNC_ROUNDS 1
#endif

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
  // one tile: stage, solve, store
  auto do_tile = [&](long long tile, long long tile_next) {
    const long long q0 = tile * 32;
    const long long rem = batch - q0;
    const int cnt = rem < 32 ? (int)rem : 32;
    LaneTileIO tio;
#pragma unroll
    for (int k = 0; k < 6; k++) {
      tio.in[k] = P.in[k] + q0 * P.stride[k];
      tio.stride[k] = P.stride[k];
    }
    if (P.ci_val) {
      tio.CIv = P.ci_val + q0 * P.s_ci_val;
      tio.CIr = P.ci_row + q0 * P.s_ci_row;
      tio.sCIv = P.s_ci_val;
      tio.sCIr = P.s_ci_row;
    }
    tio.cnt = cnt;
    if (tile_next < tiles) {  // pull the next tile's G / g0 towards L2 (phase A of the next round reads them)
      tio.next_G = P.in[0] + tile_next * 32 * P.stride[0];
      tio.next_g0 = P.in[1] + tile_next * 32 * P.stride[1];
      tio.next_CI = P.in[4] + tile_next * 32 * P.stride[4];
      tio.next_ci0 = P.in[5] + tile_next * 32 * P.stride[5];
    }
    double xo[N];
    double f;
    int status, iq, iters;
    gi_lane_tile<N, FMA>(v, ws, tio, P.max_iter, xo, f, status, iq, iters);
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
  };

#if MGQP_LANE_CTA_SYNC
  // Synchronous rounds: groups of MGQP_LANE_CTA_SYNC warps start their tiles together (a named barrier per
  // group), so that they walk through the same parts of the (large) code at about the same time and share
  // instruction-cache lines.  Measured (profiles/r02_notes.md): with eight independent warps per SM a third of
  // all stall samples were instruction-fetch misses.  The group's leader draws the tickets one round ahead, so
  // that every warp knows its next tile while it works on the current one.
  constexpr int GW = MGQP_LANE_CTA_SYNC;          // warps per group
  constexpr int SR = MGQP_LANE_SYNC_ROUNDS;       // tiles per warp between two barriers
  __shared__ unsigned s_round_base[8];
  const int grp = warp / GW, wig = warp - grp * GW;
  const int grp_warps = (warps_per_cta - grp * GW) < GW ? (warps_per_cta - grp * GW) : GW;
  auto grp_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(grp_warps * 32) : "memory"); };
  auto draw = [&](long long round) {
    if (wig == 0 && lane == 0)
      s_round_base[grp] = P.ticket ? atomicAdd(P.ticket, (unsigned)(SR * grp_warps))
                                   : (unsigned)(((round * gridDim.x + blockIdx.x) * warps_per_cta + grp * GW) * SR);
  };
  grp_sync();
  draw(0);
  grp_sync();
  long long cur = (long long)s_round_base[grp];
  for (long long round = 0; cur < tiles; round++) {
    grp_sync();
    draw(round + 1);
    grp_sync();
    const long long nxt = (long long)s_round_base[grp];
#pragma unroll 1
    for (int sub = 0; sub < SR; sub++) {  // SR tiles per warp between two barriers
      const long long tile = cur + sub * grp_warps + wig;
      const long long tile_next = (sub + 1 < SR) ? tile + grp_warps : nxt + wig;
      if (tile < tiles) do_tile(tile, tile_next);
    }
    cur = nxt;
  }
#else
  long long tile = next_tile(-1);
  while (tile < tiles) {
    const long long tile_next = next_tile(tile);
    do_tile(tile, tile_next);
    tile = tile_next;
  }
#endif
}

}  // namespace mgqp
