// Persistent warp-per-tile kernel around gi_solve_tile (n <= 32).
//
// Every WARP is an independent worker: it pulls a tile of 32/GW consecutive QPs from an atomic work
// queue, stages the tile's six input fields into its private shared-memory block with TMA bulk
// copies (cp.async.bulk, one per field, completion on the warp's own mbarrier), prefetches the
// NEXT tile into L2 while it solves the current one, and writes x/f/status straight to HBM.
// There is no __syncthreads anywhere; a CTA is just a bag of warps sized to fill the SM's shared
// memory, and the grid is one resident wave (148 SMs x CTAs/SM).
#pragma once
#include "gi_small.cuh"

namespace mgqp {

struct SmallParams {
  TileView tv;
  long long batch, tiles;
  const double* in[6];   // G, g0, CE, ce0, CI, ci0
  long long stride[6];   // element stride between QPs
  int size[6];           // elements per QP: n*n, n, n*p, p, n*m, m
  int off[6];            // offsets of the fields inside the warp block (doubles)
  unsigned bulk_mask;    // bit k: field k may be staged with one bulk copy per full tile
  unsigned bulk_bytes;   // sum of bytes of the bulk-staged fields of one full tile
  double* x; double* f; int* status; int* iq; int* A; int* iters;
  int max_iter;
  unsigned* ticket;      // work-queue head (zeroed before launch); nullptr = static round robin
  int bar_bytes;         // bytes reserved for the per-warp mbarriers at the start of shared memory
  // column-compressed CI (mgqp_batch::CI_val / CI_row) or null; when set, field 4 is not read at all
  const double* ci_val;
  const unsigned char* ci_row;
  long long s_ci_val, s_ci_row;
  const unsigned long long* batch_dev;  // if set: the number of QPs lives on the device (<= batch)
  int fuse_retry;        // TileOut::fuse_retry
  const int* gate;       // if set: run only if (*gate != 0) == (gate_want != 0) (the launcher's classification verdict)
  int gate_want;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int GW, bool FMA>
__global__ void __launch_bounds__(512, 1) gi_small_kernel(const __grid_constant__ SmallParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int QPW = 32 / GW;
  if (P.gate && (__ldg(P.gate) != 0) != (P.gate_want != 0)) return;  // the other layout serves this batch
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps_per_cta = blockDim.x >> 5;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw) + warp;
  double* ws = reinterpret_cast<double*>(smem_raw + P.bar_bytes) + (size_t)warp * P.tv.total;

  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async();
  __syncwarp();

  const long long total_warps = (long long)gridDim.x * warps_per_cta;
  const long long my_warp = (long long)blockIdx.x * warps_per_cta + warp;
  unsigned parity = 0;

  auto next_tile = [&](long long prev) -> long long {
    long long t;
    if (P.ticket) {
      unsigned v = 0;
      if (lane == 0) v = atomicAdd(P.ticket, 1u);
      t = (long long)__shfl_sync(MGQP_FULL, v, 0);
    } else {
      t = (prev < 0) ? my_warp : prev + total_warps;
    }
    return t;
  };

  long long batch = P.batch, tiles = P.tiles;
  if (P.batch_dev) {
    const long long b = (long long)*P.batch_dev;
    batch = b < batch ? b : batch;
    tiles = (batch + QPW - 1) / QPW;
  }
  long long tile = next_tile(-1);
  while (tile < tiles) {
    const long long q0 = tile * QPW;
    const long long rem = batch - q0;
    const int cnt = rem < QPW ? (int)rem : QPW;
    const unsigned bulk = (cnt == QPW) ? P.bulk_mask : 0u;

    // ---- stage the tile
    if (bulk && lane == 0) {
      mbar_expect_tx(bar, P.bulk_bytes);
#pragma unroll 1
      for (int k = 0; k < 6; k++)
        if (bulk & (1u << k))
          bulk_g2s(ws + P.off[k], P.in[k] + q0 * P.stride[k], (unsigned)(QPW * P.size[k] * 8), bar);
    }
    // fields that cannot go through TMA (odd alignment, broadcast or padded strides, tail tile):
    // plain coalesced loads; groups beyond the tail replicate the tile's first QP
    if (P.ci_val && P.size[4] > 0 && P.tv.ci_comp) {
      // column-compressed constraints stay compressed: m values and m rows (int) per QP
      const int n = P.tv.n, m = P.tv.m, cs = P.tv.cisz;
      double* ci = ws + P.off[4];
#pragma unroll 1
      for (int idx = lane; idx < QPW * m; idx += 32) {
        const int q = idx / m, i = idx - q * m;
        const long long qq = q0 + (q < cnt ? q : 0);
        unsigned r = __ldg(P.ci_row + qq * P.s_ci_row + i);
        r = r < (unsigned)n ? r : (unsigned)(n - 1);
        ci[q * cs + i] = __ldg(P.ci_val + qq * P.s_ci_val + i);
        reinterpret_cast<int*>(ci + q * cs + m)[i] = (int)r;
      }
    } else if (P.ci_val && P.size[4] > 0) {
      // column-compressed constraints: rebuild the dense n x m block in shared memory (zeros, then one
      // coefficient per column).  The kernel's sums start from +0.0, so a +0.0 where the caller's block
      // had -0.0 cannot change any result.
      const int n = P.tv.n, m = P.tv.m, sz = n * m;
      double* ci = ws + P.off[4];
#pragma unroll 1
      for (int idx = lane; idx < QPW * sz; idx += 32) ci[idx] = 0.0;
      __syncwarp();
#pragma unroll 1
      for (int idx = lane; idx < QPW * m; idx += 32) {
        const int q = idx / m, i = idx - q * m;
        const long long qq = q0 + (q < cnt ? q : 0);
        unsigned r = __ldg(P.ci_row + qq * P.s_ci_row + i);
        r = r < (unsigned)n ? r : (unsigned)(n - 1);
        ci[q * sz + (int)r * m + i] = __ldg(P.ci_val + qq * P.s_ci_val + i);
      }
    }
#pragma unroll 1
    for (int k = 0; k < 6; k++) {  // cold path: keep it small (instruction-cache footprint)
      if (k == 4 && P.ci_val) continue;
      if (!(bulk & (1u << k)) && P.size[k] > 0) {
        const int sz = P.size[k], tot = QPW * sz;
#pragma unroll 1
        for (int idx = lane; idx < tot; idx += 32) {
          const int q = idx / sz, e = idx - q * sz;
          const long long qq = q0 + (q < cnt ? q : 0);
          ws[P.off[k] + idx] = __ldg(P.in[k] + qq * P.stride[k] + e);
        }
      }
    }
    // ---- next ticket now; pull that tile towards L2 while this one is being solved
    const long long tile_next = next_tile(tile);
    if (tile_next < tiles && lane < 6 && (P.bulk_mask & (1u << lane)) &&
        (tile_next + 1) * QPW <= batch) {
      bulk_prefetch_l2(P.in[lane] + tile_next * QPW * P.stride[lane], (unsigned)(QPW * P.size[lane] * 8));
    }
    if (bulk) {
      mbar_wait(bar, parity);
      parity ^= 1u;
    }
    __syncwarp();

    TileOut out;
    out.x = P.x + q0 * P.tv.n;
    out.f = P.f + q0;
    out.status = P.status + q0;
    out.iq = P.iq ? P.iq + q0 : nullptr;
    out.A = P.A ? P.A + q0 * (P.tv.p + P.tv.m) : nullptr;
    out.iters = P.iters ? P.iters + q0 : nullptr;
    out.fuse_retry = P.fuse_retry != 0;
    gi_solve_tile<GW, FMA>(P.tv, ws, cnt, P.max_iter, out);

    // generic-proxy accesses to the block are done; order them before the next async-proxy write
    fence_proxy_async();
    __syncwarp();
    tile = tile_next;
  }
}

}  // namespace mgqp
