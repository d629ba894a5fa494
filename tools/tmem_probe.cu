// Probe: tensor memory (TMEM) as lane-private fp64 scratch -- tcgen05.alloc / st / ld from every warp
// of a CTA (warp w owns lanes 32*(w%4)..+31), correctness + latency / throughput of the 32x32b shapes.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_probe tools/tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_st2(uint32_t a, double v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(a), "r"(__double2loint(v)), "r"(__double2hiint(v)));
}
__device__ __forceinline__ double tm_ld2(uint32_t a) {
  uint32_t lo, hi;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(a));
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(lo), "+r"(hi));
  return __hiloint2double(hi, lo);
}
__device__ __forceinline__ void tm_ld2_nowait(uint32_t a, uint32_t& lo, uint32_t& hi) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(a));
}
__device__ __forceinline__ void tm_ld16_nowait(uint32_t a, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(a));
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(512, 1) probe(int ncols_per_warp, long long* out_cycles, int* out_err, int mode) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (warp == 0) {
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tbase);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * ncols_per_warp);
  const int nd = ncols_per_warp / 2;  // doubles per lane
  int err = 0;
  for (int i = 0; i < nd; i++) tm_st2(base + 2 * i, (double)(blockIdx.x * 1000003 + threadIdx.x * 131 + i) + 0.25);
  tm_wait_st();
  for (int i = 0; i < nd; i++) {
    const double v = tm_ld2(base + 2 * i);
    if (v != (double)(blockIdx.x * 1000003 + threadIdx.x * 131 + i) + 0.25) err++;
  }
  // x16 read-back of the first 8 doubles
  {
    uint32_t r[16];
    tm_ld16_nowait(base, r);
    tm_wait_ld();
    for (int i = 0; i < 8; i++) {
      const double v = __hiloint2double(r[2 * i + 1], r[2 * i]);
      if (v != (double)(blockIdx.x * 1000003 + threadIdx.x * 131 + i) + 0.25) err++;
    }
  }
  __syncwarp();
  // timing
  long long t0 = 0, t1 = 0;
  double acc = 0.0;
  const int active = (mode & 1) ? nw : 1;  // 1: all warps time concurrently, 0: warp 0 only
  __syncthreads();
  if (warp < active) {
    // (a) dependent ld.x2 + wait pairs
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; it++) acc += tm_ld2(base + 2 * (it % nd));
    t1 = clock64();
    if (lane == 0) out_cycles[(blockIdx.x * 16 + warp) * 4 + 0] = (t1 - t0) / 64;
    // (b) 8 x ld.x2 in flight, one wait
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 16; it++) {
      uint32_t lo[8], hi[8];
#pragma unroll
      for (int k = 0; k < 8; k++) tm_ld2_nowait(base + 2 * ((it * 8 + k) % nd), lo[k], hi[k]);
      tm_wait_ld();
#pragma unroll
      for (int k = 0; k < 8; k++) acc += __hiloint2double(hi[k], lo[k]);
    }
    t1 = clock64();
    if (lane == 0) out_cycles[(blockIdx.x * 16 + warp) * 4 + 1] = (t1 - t0) / 16;  // cycles per 8 doubles
    // (c) ld.x16 (8 doubles) + wait
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; it++) {
      uint32_t r[16];
      tm_ld16_nowait(base + 16 * (it % (nd / 8)), r);
      tm_wait_ld();
#pragma unroll
      for (int k = 0; k < 8; k++) acc += __hiloint2double(r[2 * k + 1], r[2 * k]);
    }
    t1 = clock64();
    if (lane == 0) out_cycles[(blockIdx.x * 16 + warp) * 4 + 2] = (t1 - t0) / 64;  // cycles per 8 doubles
    // (d) st.x2 + wait::st
    t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; it++) {
      tm_st2(base + 2 * (it % nd), acc);
      tm_wait_st();
    }
    t1 = clock64();
    if (lane == 0) out_cycles[(blockIdx.x * 16 + warp) * 4 + 3] = (t1 - t0) / 64;
  }
  if (acc == 123.456) err += 1000;
  atomicAdd(out_err, err);
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

int main() {
  long long* cyc;
  int* err;
  cudaMalloc(&cyc, 148 * 16 * 4 * 8);
  cudaMalloc(&err, 4);
  for (int nw : {4, 8, 12, 16}) {
    for (int mode : {0, 1}) {
      cudaMemset(cyc, 0, 148 * 16 * 4 * 8);
      cudaMemset(err, 0, 4);
      const int ncols = 512 / ((nw + 3) / 4) & ~15;
      probe<<<148, nw * 32>>>(ncols, cyc, err, mode);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[16 * 4];
      int herr = -1;
      cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
      cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
      printf("warps=%2d cols/warp=%3d mode=%s: %s, mismatches=%d | warp0: ld2+wait %lld cyc, 8xld2+wait %lld cyc, ld16+wait %lld cyc, st2+wait %lld cyc",
             nw, ncols, mode ? "all warps" : "warp 0  ", cudaGetErrorString(e), herr, h[0], h[1], h[2], h[3]);
      if (mode) printf(" | last warp: %lld %lld %lld %lld", h[(nw - 1) * 4], h[(nw - 1) * 4 + 1], h[(nw - 1) * 4 + 2], h[(nw - 1) * 4 + 3]);
      printf("\n");
    }
  }
  return 0;
}
