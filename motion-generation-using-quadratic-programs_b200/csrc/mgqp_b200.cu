// C-ABI of the B200 batched QP engine (declared in include/mgqp_b200.h).
//
// Replaces, for batches, the reference's only solver call
//   sum = solve_quadprog(G, g0, t(CE), ce0, t(CI), ci0, x);          src/mgqp.cpp:708, :725
// (prototype include/QuadProgpp/QuadProg++.hh:69-72, body in lib/QuadProgpp/libquadprog.a).
// There is no CPU fallback in this file: without a CUDA device every entry point fails loudly.
#include "../../include/mgqp_b200.h"

#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "gi_lane_kernel.cuh"
#include "gi_large_kernel.cuh"
#include "gi_small_kernel.cuh"
#include "retry.cuh"
#include "cascade.cuh"

namespace {

using namespace mgqp;

thread_local std::string g_last_error;
thread_local long long g_h2d_bytes = 0, g_d2h_bytes = 0;  // of the calling thread's last host-buffer call
thread_local std::vector<long long> g_multi_chunks;       // chunks each device of the last multi-device call took
std::atomic<long long> g_launches{0};

int fail(int rc, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return rc;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(MGQP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                  __LINE__);                                                                       \
  } while (0)

constexpr int kMaxDev = 16;
// internal flag bit (never part of the ABI): the first solve of a retry call may answer the retry itself where
// that is exact (gi_small.cuh, TileOut::fuse_retry)
constexpr int kFlagFuseRetry = 1 << 30;
// internal: the caller knows its dense inequality block is not made of bound columns (the cascade's projected
// limits): skip the classification pre-pass of the group kernel
constexpr int kFlagDenseCI = 1 << 29;
// Work-queue heads, one slot per launch, reused round robin: two launches only ever share a slot if
// more than kTicketRing launches of this library are in flight on one device at the same time.
constexpr int kTicketRing = 4096;

struct DeviceCtx {
  std::mutex mu;
  bool init = false;
  int sms = 0;
  int smem_optin = 0;
  unsigned* tickets = nullptr;  // ring of work-queue heads
  unsigned ring_pos = 0;
  bool attr_set[3][2] = {{false, false}, {false, false}, {false, false}};
  int attr_smem[3][2] = {{0, 0}, {0, 0}, {0, 0}};
  int lane_attr_smem[9][2] = {};
  int large_attr_smem[2] = {0, 0};
  // host-buffer pipeline: two streams and two staging arenas, kept across calls (grow only)
  std::mutex host_mu;
  cudaStream_t host_st[2] = {nullptr, nullptr};
  cudaEvent_t host_up[2] = {nullptr, nullptr};  // the uploads of the buffer set's current chunk are done
  void* arena[2] = {nullptr, nullptr};
  size_t arena_cap[2] = {0, 0};
  // pinned host staging for the column-compressed CI of one chunk (values, rows), per buffer set
  void* pack_host[2] = {nullptr, nullptr};
  size_t pack_cap[2] = {0, 0};
  // the library's own pool for its stream-ordered workspaces (CTA-per-QP kernel, retry pass): the device's
  // default pool belongs to the host application
  cudaMemPool_t pool = nullptr;
  // single-QP path (mgqp_solve_one, the C++ drop-in): one stream, one pinned and one device staging block
  std::mutex one_mu;
  cudaStream_t one_st = nullptr;
  double* one_host = nullptr;
  double* one_dev = nullptr;
  size_t one_cap = 0;  // doubles
  // cascade: G = I14, g0 = 0 and limitsMatrix = [-I; +I] in column-compressed form, one instance for every scenario
  double* casc_const = nullptr;
  cudaStream_t casc_st[2] = {nullptr, nullptr};
  void* casc_arena[2] = {nullptr, nullptr};
  size_t casc_cap[2] = {0, 0};
};
DeviceCtx g_dev[kMaxDev];

int device_ctx(int device, DeviceCtx** out) {
  if (device < 0 || device >= kMaxDev) return fail(MGQP_ERR_INVALID_ARG, "device %d out of range", device);
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return fail(MGQP_ERR_NO_DEVICE, "no CUDA device available: this engine has no CPU fallback");
  if (device >= count) return fail(MGQP_ERR_INVALID_ARG, "device %d >= device count %d", device, count);
  DeviceCtx& c = g_dev[device];
  std::lock_guard<std::mutex> lk(c.mu);
  if (!c.init) {
    CU(cudaSetDevice(device));
    CU(cudaDeviceGetAttribute(&c.sms, cudaDevAttrMultiProcessorCount, device));
    CU(cudaDeviceGetAttribute(&c.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    {
      // stream-ordered workspaces (CTA-per-QP kernel, retry pass) are allocated per call from a pool of the
      // library's own: freed blocks stay cached there (up to 6 GB -- the CTA-per-QP kernel's workspace for the
      // horizon-stacked shape is 1.3 GB, a cascade chunk 1.5 GB; handing those back to the driver at every
      // synchronisation made consecutive calls vary by 4x -- more goes back at the next synchronisation) instead of
      // being parked in the device's default pool, which the application uses
      cudaMemPoolProps props;
      memset(&props, 0, sizeof props);
      props.allocType = cudaMemAllocationTypePinned;
      props.handleTypes = cudaMemHandleTypeNone;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = device;
      CU(cudaMemPoolCreate(&c.pool, &props));
      unsigned long long keep = 6ull << 30;
      cudaMemPoolSetAttribute(c.pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    CU(cudaMalloc(&c.tickets, kTicketRing * sizeof(unsigned)));
    CU(cudaMemset(c.tickets, 0, kTicketRing * sizeof(unsigned)));
    c.init = true;
  }
  *out = &c;
  return MGQP_OK;
}

struct Plan {
  int gw = -1;         // lanes per QP (8/16/32); 0 = CTA-per-QP kernel; <0 unsupported
  int warps = 0;       // warps per CTA
  int smem = 0;        // dynamic shared memory per CTA
  int bar_bytes = 0;
  TileView tv;
};

// SM shared-memory budget used for planning when no device is queried (B200: 227 KB opt-in).
constexpr int kSmemBudget = 227 * 1024;

#ifndef MGQP_SMALL_CI_COMP
#define MGQP_SMALL_CI_COMP 1   // column-compressed CI stays compressed in the group kernel's shared memory
#endif
Plan make_plan(int n, int p, int m, int smem_optin, int ci_comp = 0) {
  Plan pl;
  if (n < 1 || p < 0 || m < 0) return pl;
  int gw = n <= 8 ? 8 : n <= 16 ? 16 : n <= 32 ? 32 : 0;
  while (gw && gw < 32 && m > 32 * gw) gw *= 2;
  if (gw == 0 || m > 32 * gw) {
    pl.gw = -2;
    return pl;
  }
  for (;; gw *= 2) {
    TileView tv = make_tile_view(n, p, m, gw, ci_comp);
    const long long per_warp = (long long)tv.total * 8;
    // one CTA per SM unless warps are small; leave the 1 KB/CTA the driver reserves
    int warps = (int)((smem_optin - 1024 - 256) / per_warp);
    if (warps >= 1) {
      if (warps > 16) warps = 16;
      pl.gw = gw;
      pl.warps = warps;
      pl.tv = tv;
      pl.bar_bytes = ((warps * 8) + 15) & ~15;
      pl.smem = pl.bar_bytes + (int)(per_warp * warps);
      return pl;
    }
    if (gw == 32) break;
  }
  pl.gw = -2;
  return pl;
}

template <int GW, bool FMA>
int launch_small(DeviceCtx* ctx, const Plan& pl, const SmallParams& P, cudaStream_t st, int gi, int fi) {
  auto kern = gi_small_kernel<GW, FMA>;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->attr_set[gi][fi] || ctx->attr_smem[gi][fi] < pl.smem) {
      CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem));
      ctx->attr_set[gi][fi] = true;
      ctx->attr_smem[gi][fi] = pl.smem;
    }
  }
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, pl.warps * 32, pl.smem));
  if (per_sm < 1) return fail(MGQP_ERR_UNSUPPORTED_SHAPE, "kernel does not fit an SM (smem %d B)", pl.smem);
  long long ctas = (long long)ctx->sms * per_sm;
  const long long need = (P.tiles + pl.warps - 1) / pl.warps;
  if (ctas > need) ctas = need;
  if (ctas < 1) ctas = 1;
  kern<<<(unsigned)ctas, pl.warps * 32, pl.smem, st>>>(P);
  g_launches.fetch_add(1);
  CU(cudaGetLastError());
  return MGQP_OK;
}


// ---------------------------------------------------------------- lane-per-QP family (n <= 8)
struct LanePlan {
  bool ok = false;
  int warps = 0, smem = 0;
  LaneView v;
};

LanePlan make_lane_plan(int n, int p, int m, int smem_optin) {
  LanePlan lp;
  if (n < 1 || n > 8 || m > 64 || p > n) return lp;
  lp.v = make_lane_view(n, p, m);
  const long long per_warp = 32ll * lp.v.stride * 8;
  int warps = (int)((smem_optin - 1024) / per_warp);
  if (warps < 2) return lp;  // one warp per SM cannot hide anything: use the lane-group kernel instead
  if (warps > 8) warps = 8;
  if (const char* e = std::getenv("MGQP_LANE_WARPS")) {  // tuning knob: fewer resident warps per SM
    const int w = std::atoi(e);
    if (w >= 1 && w < warps) warps = w;
  }
  lp.warps = warps;
  lp.smem = (int)(per_warp * warps);
  lp.ok = true;
  return lp;
}

template <int N, bool FMA>
int launch_lane(DeviceCtx* ctx, const LanePlan& lp, const LaneParams& P, cudaStream_t st) {
  auto kern = gi_lane_kernel<N, FMA>;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->lane_attr_smem[N][FMA] < lp.smem) {
      CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lp.smem));
      ctx->lane_attr_smem[N][FMA] = lp.smem;
    }
  }
  if (const char* e = std::getenv("MGQP_LANE_CARVEOUT"))  // tuning knob: shared-memory carve-out in percent
    CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, std::atoi(e)));
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, lp.warps * 32, lp.smem));
  if (per_sm < 1) return fail(MGQP_ERR_UNSUPPORTED_SHAPE, "lane kernel does not fit an SM (smem %d B)", lp.smem);
  if (const char* e = std::getenv("MGQP_LANE_CTAS_PER_SM")) {  // tuning knob: fewer resident CTAs per SM
    const int c = std::atoi(e);
    if (c >= 1 && c < per_sm) per_sm = c;
  }
  long long ctas = (long long)ctx->sms * per_sm;
  const long long need = (P.tiles + lp.warps - 1) / lp.warps;
  if (ctas > need) ctas = need;
  if (ctas < 1) ctas = 1;
  kern<<<(unsigned)ctas, lp.warps * 32, lp.smem, st>>>(P);
  g_launches.fetch_add(1);
  CU(cudaGetLastError());
  return MGQP_OK;
}

template <bool FMA>
int launch_lane_n(DeviceCtx* ctx, const LanePlan& lp, const LaneParams& P, cudaStream_t st) {
  switch (lp.v.n) {
    case 1: return launch_lane<1, FMA>(ctx, lp, P, st);
    case 2: return launch_lane<2, FMA>(ctx, lp, P, st);
    case 3: return launch_lane<3, FMA>(ctx, lp, P, st);
    case 4: return launch_lane<4, FMA>(ctx, lp, P, st);
    case 5: return launch_lane<5, FMA>(ctx, lp, P, st);
    case 6: return launch_lane<6, FMA>(ctx, lp, P, st);
    case 7: return launch_lane<7, FMA>(ctx, lp, P, st);
    default: return launch_lane<8, FMA>(ctx, lp, P, st);
  }
}


// ---------------------------------------------------------------- CTA-per-QP family (any n)
struct LargePlan {
  bool ok = false;
  int T = 0;
  LargeView v;
};

LargePlan make_large_plan(int n, int p, int m, int smem_optin) {
  LargePlan lp;
  int want = n > (m + 3) / 4 ? n : (m + 3) / 4;
  int T = 64;
  int tmax = 128;  // (gi_large_kernel.cuh: launch bounds 128 x 5 -- small CTAs, many per SM)
  if (const char* e = std::getenv("MGQP_LARGE_T")) tmax = std::atoi(e);  // tuning knob
  if (tmax > MGQP_LARGE_BOUNDS_T) tmax = MGQP_LARGE_BOUNDS_T;
  while (T < want && T < tmax) T *= 2;
  if (const char* e = std::getenv("MGQP_LARGE_T_EXACT")) {  // tuning knob: any multiple of 32
    const int t = std::atoi(e);
    if (t >= 32 && t <= MGQP_LARGE_BOUNDS_T && t % 32 == 0) T = t;
  }
  lp.T = T;
  lp.v = make_large_view(n, p, m, T);
  lp.ok = lp.v.smem_bytes + 1024 <= smem_optin && (long long)n * n < (1ll << 30);
  return lp;
}

template <bool FMA>
int launch_large(DeviceCtx* ctx, const LargePlan& lp, LargeParams& P, cudaStream_t st) {
  auto kern = gi_large_kernel<FMA>;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->large_attr_smem[FMA] < lp.v.smem_bytes) {
      CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, lp.v.smem_bytes));
      ctx->large_attr_smem[FMA] = lp.v.smem_bytes;
    }
  }
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, lp.T, lp.v.smem_bytes));
  if (per_sm < 1) return fail(MGQP_ERR_UNSUPPORTED_SHAPE, "CTA-per-QP kernel does not fit an SM (smem %d B)", lp.v.smem_bytes);
  int cap_sm = MGQP_LARGE_BOUNDS_B;
  if (const char* e = std::getenv("MGQP_LARGE_CTAS_PER_SM")) cap_sm = std::atoi(e);  // tuning knob
  if (per_sm > cap_sm) per_sm = cap_sm;
  long long ctas = (long long)ctx->sms * per_sm;
  if (const char* e = std::getenv("MGQP_LARGE_MAX_CTAS")) {  // tuning knob: fewer CTAs in flight (L2 residency)
    const long long c = std::atoll(e);
    if (c >= 1 && c < ctas) ctas = c;
  }
  if (ctas > P.batch) ctas = P.batch;
  const long long ws_bytes = lp.v.ws_doubles * 8;
  const long long cap = (8ll << 30) / ws_bytes;  // keep the workspace under 8 GB
  if (ctas > cap) ctas = cap;
  if (ctas < 1) return fail(MGQP_ERR_UNSUPPORTED_SHAPE, "workspace of %lld bytes per QP is too large", ws_bytes);
  void* ws = nullptr;
  CU(cudaMallocFromPoolAsync(&ws, (size_t)(ctas * ws_bytes), ctx->pool, st));
  P.workspace = static_cast<double*>(ws);
  // set-up admission gate (gi_large.cuh): off by default -- measured on C4 (profiles/r02_notes.md): limiting the
  // number of CTAs that factorise at once to what fits L2 costs more in waiting than it saves in DRAM latency
  {
    long long limit = 0;
    if (const char* e = std::getenv("MGQP_LARGE_SETUP_GATE")) limit = std::atoll(e);  // tuning knob (0 = off)
    if (limit >= 1 && limit < ctas) {
      unsigned slot;
      {
        std::lock_guard<std::mutex> lk(ctx->mu);
        slot = ctx->ring_pos++ % kTicketRing;
      }
      P.gate = ctx->tickets + slot;
      P.gate_limit = (int)limit;
      CU(cudaMemsetAsync(P.gate, 0, sizeof(unsigned), st));
    }
  }
  kern<<<(unsigned)ctas, lp.T, lp.v.smem_bytes, st>>>(P);
  g_launches.fetch_add(1);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(ws, st);  // stream-ordered: released after the kernel has run
  if (e != cudaSuccess) return fail(MGQP_ERR_CUDA, "gi_large_kernel launch failed: %s", cudaGetErrorString(e));
  return MGQP_OK;
}

// the caller's current device is put back when an entry point returns
// tuning / A-B knob: MGQP_SMALL_CLASSIFY=0 keeps a dense block dense (the behaviour before the pre-pass existed)
bool classify_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("MGQP_SMALL_CLASSIFY");
    return !(e && e[0] == '0');
  }();
  return on;
}

// Device-side counterpart of mgqp_pack_bound_columns (host_pack.cc) for a DENSE inequality block handed to the
// device-resident entry point: one thread per (QP, column) walks the column's n coefficients (+-0.0 are the zeros, NaN
// counts as a nonzero) and writes the (value, row) pair -- (+0.0, row 0) for an all-zero column -- or raises *mixed
// if the column has a second nonzero.  The group kernel is then launched twice, gated on *mixed: the compressed layout
// (more warps per SM) runs if every column of the batch was a bound, the dense layout otherwise; the other launch
// exits at once.  No host synchronisation: the call stays asynchronous and graph-capturable.
__global__ void __launch_bounds__(256) bound_columns_kernel(long long total, int n, int m, const double* __restrict__ CI,
                                                            long long sCI, double* __restrict__ val,
                                                            unsigned char* __restrict__ row, int* mixed) {
  // One look at the verdict per block and one write per block that finds a second nonzero: every thread polling or
  // raising the same word cost 0.3 ms per launch whatever the batch (`r06g`).
  __shared__ int stop;
  if (threadIdx.x == 0) stop = *reinterpret_cast<volatile int*>(mixed);
  __syncthreads();
  if (stop) return;  // verdict known: the compressed launch will not run
  const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
  double v = 0.0;
  int r = 0, cnt = 0;
  if (e < total) {
    const long long q = e / m;
    const int i = (int)(e - q * m);
    const double* c = CI + q * sCI + i;
    int j = 0;
    for (; j + 4 <= n; j += 4) {  // four loads in flight per thread
      const double a0 = __ldg(c + (long long)j * m), a1 = __ldg(c + (long long)(j + 1) * m);
      const double a2 = __ldg(c + (long long)(j + 2) * m), a3 = __ldg(c + (long long)(j + 3) * m);
      if ((__double_as_longlong(a0) & 0x7fffffffffffffffll) != 0) { v = a0; r = j; cnt++; }
      if ((__double_as_longlong(a1) & 0x7fffffffffffffffll) != 0) { v = a1; r = j + 1; cnt++; }
      if ((__double_as_longlong(a2) & 0x7fffffffffffffffll) != 0) { v = a2; r = j + 2; cnt++; }
      if ((__double_as_longlong(a3) & 0x7fffffffffffffffll) != 0) { v = a3; r = j + 3; cnt++; }
    }
    for (; j < n; j++) {
      const double a = __ldg(c + (long long)j * m);
      if ((__double_as_longlong(a) & 0x7fffffffffffffffll) != 0) { v = a; r = j; cnt++; }
    }
  }
  if (__syncthreads_or(cnt > 1)) {
    if (threadIdx.x == 0) *mixed = 1;
    return;
  }
  if (e < total) {
    val[e] = v;
    row[e] = (unsigned char)r;
  }
}

struct DeviceGuard {
  int prev = -1;
  DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int validate(const mgqp_batch* d) {
  if (!d) return fail(MGQP_ERR_INVALID_ARG, "null batch descriptor");
  if (d->n < 1) return fail(MGQP_ERR_INVALID_ARG, "n = %d must be >= 1", d->n);
  if (d->p < 0 || d->m < 0) return fail(MGQP_ERR_INVALID_ARG, "p = %d, m = %d must be >= 0", d->p, d->m);
  if (d->batch < 0) return fail(MGQP_ERR_INVALID_ARG, "batch = %lld must be >= 0", (long long)d->batch);
  if (!d->G || !d->g0 || (d->p && (!d->CE || !d->ce0)) || (d->m && ((!d->CI && !d->CI_val) || !d->ci0)))
    return fail(MGQP_ERR_INVALID_ARG, "null input pointer");
  if (d->CI_val && d->m && !d->CI_row) return fail(MGQP_ERR_INVALID_ARG, "CI_val without CI_row");
  if (d->CI_val && (d->sCIv < 0 || d->sCIr < 0 || (d->sCIv && d->sCIv < d->m) || (d->sCIr && d->sCIr < d->m)))
    return fail(MGQP_ERR_INVALID_ARG, "bad stride of the column-compressed CI");
  if (!d->x || !d->f || !d->status) return fail(MGQP_ERR_INVALID_ARG, "null output pointer (x, f, status are required)");
  if (d->p > d->n)
    return fail(MGQP_ERR_INVALID_ARG, "p = %d equalities on n = %d variables (the reference indexes R out of bounds there)",
                d->p, d->n);
  if (d->sG < 0 || d->sg0 < 0 || d->sCE < 0 || d->sce0 < 0 || d->sCI < 0 || d->sci0 < 0)
    return fail(MGQP_ERR_INVALID_ARG, "negative stride");
  return MGQP_OK;
}

int solve_device_once(const mgqp_batch* d, int device, cudaStream_t st, const unsigned long long* batch_dev = nullptr,
                      const int* gate = nullptr, int gate_want = 0) {
  int rc = validate(d);
  if (rc) return rc;
  if (d->batch == 0) return MGQP_OK;
  DeviceCtx* ctx = nullptr;
  if ((rc = device_ctx(device, &ctx))) return rc;
  CU(cudaSetDevice(device));
  const int max_iter = d->max_iter > 0 ? d->max_iter : 100 + 10 * (d->p + d->m);
  unsigned* ticket = nullptr;
  if (!(d->flags & MGQP_FLAG_STATIC_SCHEDULE)) {
    unsigned slot;
    {
      std::lock_guard<std::mutex> lk(ctx->mu);
      slot = ctx->ring_pos++ % kTicketRing;
    }
    ticket = ctx->tickets + slot;
    CU(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
  }
  const bool fast = (d->flags & MGQP_FLAG_FAST_MATH) != 0;
  if (!(d->flags & (MGQP_FLAG_GROUP_KERNEL | MGQP_FLAG_CTA_KERNEL))) {
    const LanePlan lp = make_lane_plan(d->n, d->p, d->m, ctx->smem_optin);
    if (lp.ok) {
      LaneParams L;
      memset(&L, 0, sizeof L);
      L.v = lp.v;
      L.batch = d->batch;
      L.tiles = (d->batch + 31) / 32;
      const double* in[6] = {d->G, d->g0, d->CE, d->ce0, d->CI, d->ci0};
      const long long stride[6] = {d->sG, d->sg0, d->sCE, d->sce0, d->sCI, d->sci0};
      for (int k = 0; k < 6; k++) {
        L.in[k] = in[k];
        L.stride[k] = stride[k];
      }
      L.x = d->x; L.f = d->f; L.status = d->status; L.iq = d->iq; L.A = d->A; L.iters = d->iters;
      L.max_iter = max_iter;
      L.ticket = ticket;
      L.batch_dev = batch_dev;
      if (d->CI_val && d->m) {
        L.ci_val = d->CI_val;
        L.ci_row = d->CI_row;
        L.s_ci_val = d->sCIv;
        L.s_ci_row = d->sCIr;
      }
      return fast ? launch_lane_n<true>(ctx, lp, L, st) : launch_lane_n<false>(ctx, lp, L, st);
    }
  }

  const Plan pl = (d->flags & MGQP_FLAG_CTA_KERNEL)
                      ? Plan()
                      : make_plan(d->n, d->p, d->m, ctx->smem_optin, (MGQP_SMALL_CI_COMP && d->CI_val && d->m > 0) ? 1 : 0);
  // A dense inequality block whose columns may all be bounds (every one the reference builds is): classify it on
  // the device and let the layout with more resident warps serve the batch if it qualifies.
  if (MGQP_SMALL_CI_COMP && pl.gw > 0 && !gate && !batch_dev && d->CI && !d->CI_val && d->m > 0 && d->n <= 255 &&
      !(d->flags & (kFlagDenseCI | MGQP_FLAG_NO_HOST_PACK)) && classify_enabled()) {
    const Plan pc = make_plan(d->n, d->p, d->m, ctx->smem_optin, 1);
    // more warps only pay once the batch outgrows one resident wave of the dense layout
    if (pc.gw == pl.gw && pc.warps > pl.warps && d->batch > (long long)ctx->sms * pl.warps * pl.tv.qpw) {
      const long long cnt = d->sCI ? d->batch : 1, m = d->m;
      const size_t val_bytes = (size_t)(cnt * m) * 8, row_bytes = ((size_t)(cnt * m) + 15) & ~(size_t)15;
      void* buf = nullptr;
      CU(cudaMallocFromPoolAsync(&buf, val_bytes + row_bytes + 16, ctx->pool, st));
      double* val = static_cast<double*>(buf);
      unsigned char* row = static_cast<unsigned char*>(buf) + val_bytes;
      int* mixed = reinterpret_cast<int*>(row + row_bytes);
      int rc2 = MGQP_OK;
      cudaError_t e = cudaMemsetAsync(mixed, 0, sizeof(int), st);
      if (e == cudaSuccess) {
        const long long blocks = (cnt * m + 255) / 256;
        bound_columns_kernel<<<(unsigned)blocks, 256, 0, st>>>(cnt * m, d->n, d->m, d->CI, d->sCI, val, row, mixed);
        g_launches.fetch_add(1);
        e = cudaGetLastError();
      }
      if (e != cudaSuccess) rc2 = fail(MGQP_ERR_CUDA, "bound_columns_kernel: %s", cudaGetErrorString(e));
      if (!rc2) {
        mgqp_batch c = *d;
        c.CI_val = val; c.CI_row = row;
        c.sCIv = d->sCI ? m : 0; c.sCIr = d->sCI ? m : 0;
        rc2 = solve_device_once(&c, device, st, nullptr, mixed, 0);
      }
      if (!rc2) rc2 = solve_device_once(d, device, st, nullptr, mixed, 1);
      cudaFreeAsync(buf, st);
      return rc2;
    }
  }
  if (pl.gw < 0) {
    if (d->CI_val && d->m)
      return fail(MGQP_ERR_UNSUPPORTED_SHAPE,
                  "column-compressed CI is served by the lane and lane-group kernels only (n <= 32, no CTA-kernel flag)");
    const LargePlan lp = make_large_plan(d->n, d->p, d->m, ctx->smem_optin);
    if (!lp.ok)
      return fail(MGQP_ERR_UNSUPPORTED_SHAPE, "shape n=%d p=%d m=%d needs %d B of shared memory per CTA", d->n, d->p,
                  d->m, lp.v.smem_bytes);
    LargeParams L;
    memset(&L, 0, sizeof L);
    L.v = lp.v;
    L.batch = d->batch;
    const double* lin[6] = {d->G, d->g0, d->CE, d->ce0, d->CI, d->ci0};
    const long long lstride[6] = {d->sG, d->sg0, d->sCE, d->sce0, d->sCI, d->sci0};
    for (int k = 0; k < 6; k++) {
      L.in[k] = lin[k];
      L.stride[k] = lstride[k];
    }
    L.x = d->x; L.f = d->f; L.status = d->status; L.iq = d->iq; L.A = d->A; L.iters = d->iters;
    L.max_iter = max_iter;
    L.ticket = ticket;
    L.batch_dev = batch_dev;
    return fast ? launch_large<true>(ctx, lp, L, st) : launch_large<false>(ctx, lp, L, st);
  }

  SmallParams P;
  memset(&P, 0, sizeof P);
  P.tv = pl.tv;
  P.batch = d->batch;
  const int qpw = pl.tv.qpw;
  P.tiles = (d->batch + qpw - 1) / qpw;
  const double* in[6] = {d->G, d->g0, d->CE, d->ce0, d->CI, d->ci0};
  const long long stride[6] = {d->sG, d->sg0, d->sCE, d->sce0, d->sCI, d->sci0};
  const int size[6] = {d->n * d->n, d->n, d->n * d->p, d->p, d->n * d->m, d->m};
  const int off[6] = {pl.tv.oG, pl.tv.og0, pl.tv.oCE, pl.tv.oce0, pl.tv.oCI, pl.tv.oci0};
  P.bulk_mask = 0;
  P.bulk_bytes = 0;
  for (int k = 0; k < 6; k++) {
    P.in[k] = in[k];
    P.stride[k] = stride[k];
    P.size[k] = size[k];
    P.off[k] = off[k];
    const long long bytes = (long long)qpw * size[k] * 8;
    const bool ok = !(d->flags & MGQP_FLAG_NO_BULK_COPY) && size[k] > 0 && stride[k] == size[k] &&
                    (bytes % 16) == 0 && (reinterpret_cast<uintptr_t>(in[k]) % 16) == 0 &&
                    !(k == 4 && d->CI_val);
    if (ok) {
      P.bulk_mask |= 1u << k;
      P.bulk_bytes += (unsigned)bytes;
    }
  }
  P.x = d->x;
  P.f = d->f;
  P.status = d->status;
  P.iq = d->iq;
  P.A = d->A;
  P.iters = d->iters;
  P.max_iter = max_iter;
  if (d->CI_val && d->m) {
    P.ci_val = d->CI_val;
    P.ci_row = d->CI_row;
    P.s_ci_val = d->sCIv;
    P.s_ci_row = d->sCIr;
  }
  P.bar_bytes = pl.bar_bytes;
  P.ticket = ticket;
  P.batch_dev = batch_dev;
  P.fuse_retry = (d->flags & kFlagFuseRetry) ? 1 : 0;
  P.gate = gate;
  P.gate_want = gate_want;
  const int gi = pl.gw == 8 ? 0 : pl.gw == 16 ? 1 : 2;
  if (pl.gw == 8) return fast ? launch_small<8, true>(ctx, pl, P, st, gi, 1) : launch_small<8, false>(ctx, pl, P, st, gi, 0);
  if (pl.gw == 16) return fast ? launch_small<16, true>(ctx, pl, P, st, gi, 1) : launch_small<16, false>(ctx, pl, P, st, gi, 0);
  return fast ? launch_small<32, true>(ctx, pl, P, st, gi, 1) : launch_small<32, false>(ctx, pl, P, st, gi, 0);
}

// stream-ordered allocation released when the scope ends, whichever way it ends
struct AsyncBuf {
  void* p = nullptr;
  cudaStream_t st = nullptr;
  ~AsyncBuf() {
    if (p) cudaFreeAsync(p, st);
  }
};

// The reference adapter's policy for an infeasible answer (src/mgqp.cpp:717-735), for a whole batch:
// see retry.cuh.  Asynchronous like the plain call: the second launch reads the number of QPs that go
// round again from device memory.
int solve_device(const mgqp_batch* d, int device, cudaStream_t st) {
  if (!d) return solve_device_once(d, device, st);
  if (!(d->flags & MGQP_FLAG_RETRY_WITHOUT_CI)) {
    if (!(d->flags & kFlagFuseRetry)) return solve_device_once(d, device, st);
    mgqp_batch plain = *d;  // the internal bit is not part of the ABI
    plain.flags &= ~kFlagFuseRetry;
    return solve_device_once(&plain, device, st);
  }
  mgqp_batch first = *d;
  first.flags &= ~MGQP_FLAG_RETRY_WITHOUT_CI;
  first.flags |= kFlagFuseRetry;
  int rc = solve_device_once(&first, device, st);
  if (rc || d->batch == 0) return rc;
  if (d->n > 1024) return fail(MGQP_ERR_UNSUPPORTED_SHAPE, "MGQP_FLAG_RETRY_WITHOUT_CI serves n <= 1024");
  const long long B = d->batch, n = d->n, p = d->p, m = d->m;
  DeviceCtx* rctx = nullptr;
  if ((rc = device_ctx(device, &rctx))) return rc;
  AsyncBuf idx_buf, ws_buf;
  idx_buf.st = ws_buf.st = st;
  CU(cudaMallocFromPoolAsync(&idx_buf.p, (size_t)B * 8 + 8, rctx->pool, st));
  long long* idx = static_cast<long long*>(idx_buf.p);
  unsigned long long* count = reinterpret_cast<unsigned long long*>(idx + B);
  CU(cudaMemsetAsync(count, 0, 8, st));
  retry_mark_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(d->f, d->status, B, idx, count);
  g_launches.fetch_add(1);
  // Size the compact second batch for the worst case (every QP goes round again) and let the kernels read
  // the real count from the device, so that the call stays asynchronous; only when that workspace would be
  // unreasonably large does the host wait for the count.
  const size_t per_qp = (size_t)(n * n + n + n * p + p + n + 1) * 8 + (size_t)(3 + p) * 4;
  unsigned long long cnt = (unsigned long long)B;
  const bool on_device_count = per_qp * (size_t)B <= ((size_t)2 << 30);
  if (!on_device_count) {
    CU(cudaMemcpyAsync(&cnt, count, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (cnt == 0) return MGQP_OK;
  }
  // compact second batch: G' g0 CE ce0 | x f | status iq A iters
  const long long c = (long long)cnt;
  const size_t dbl = (size_t)c * (n * n + n + n * p + p + n + 1);
  const size_t ints = (size_t)c * (3 + p);
  CU(cudaMallocFromPoolAsync(&ws_buf.p, dbl * 8 + ints * 4 + 64, rctx->pool, st));
  double* G2 = static_cast<double*>(ws_buf.p);
  double* g02 = G2 + c * n * n;
  double* CE2 = g02 + c * n;
  double* ce02 = CE2 + c * n * p;
  double* x2 = ce02 + c * p;
  double* f2 = x2 + c * n;
  int* st2 = reinterpret_cast<int*>(f2 + c);
  int* iq2 = st2 + c;
  int* it2 = iq2 + c;
  int* A2 = it2 + c;
  RetryGather g;
  g.n = (int)n; g.p = (int)p;
  g.G = d->G; g.g0 = d->g0; g.CE = d->CE; g.ce0 = d->ce0;
  g.sG = d->sG; g.sg0 = d->sg0; g.sCE = d->sCE; g.sce0 = d->sce0;
  g.G2 = G2; g.g02 = g02; g.CE2 = CE2; g.ce02 = ce02;
  g.idx = idx; g.count = count;
  {
    const long long warps = c < 4096 ? c : 4096;
    retry_gather_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, st>>>(g);
    g_launches.fetch_add(1);
  }
  mgqp_batch second;
  memset(&second, 0, sizeof second);
  second.n = (int32_t)n; second.p = (int32_t)p; second.m = 0; second.batch = c;
  second.G = G2; second.g0 = g02; second.CE = CE2; second.ce0 = ce02;
  second.sG = n * n; second.sg0 = n; second.sCE = n * p; second.sce0 = p;
  second.x = x2; second.f = f2; second.status = st2; second.iq = iq2; second.A = p ? A2 : nullptr; second.iters = it2;
  second.max_iter = d->max_iter;
  second.flags = first.flags & ~kFlagFuseRetry;
  rc = solve_device_once(&second, device, st, count);
  if (!rc) {
    RetryScatter sc;
    sc.n = (int)n; sc.pm = (int)(p + m); sc.p = (int)p;
    sc.x2 = x2; sc.f2 = f2; sc.st2 = st2; sc.iq2 = iq2; sc.A2 = A2; sc.it2 = it2;
    sc.x = d->x; sc.f = d->f; sc.status = d->status; sc.iq = d->iq; sc.A = d->A; sc.iters = d->iters;
    sc.idx = idx; sc.count = count;
    retry_scatter_kernel<<<(unsigned)((c + 127) / 128), 128, 0, st>>>(sc);
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = fail(MGQP_ERR_CUDA, "retry kernels failed: %s", cudaGetErrorString(e));
  }
  return rc;
}

// ------------------------------------------------------------------ host-buffer pipeline
struct DevBufs {
  double *G = nullptr, *g0 = nullptr, *CE = nullptr, *ce0 = nullptr, *CI = nullptr, *ci0 = nullptr, *x = nullptr, *f = nullptr;
  int *status = nullptr, *iq = nullptr, *A = nullptr, *iters = nullptr;
  double* CIv = nullptr;          // column-compressed CI of the chunk
  unsigned char* CIr = nullptr;
  double* hCIv = nullptr;         // its pinned host staging (packer output)
  unsigned char* hCIr = nullptr;
};

// rows of the compressed form are padded so that every QP's block stays 16-byte aligned
inline long long packed_row_stride(long long m) { return (m + 15) & ~15ll; }

// Carve the staging buffers of one chunk out of one arena (256-byte aligned pieces).
size_t carve_bufs(DevBufs& b, const mgqp_batch* h, long long chunk, char* base) {
  const long long n = h->n, p = h->p, m = h->m;
  size_t off = 0;
  auto take = [&](long long bytes) -> char* {
    char* ptr = base ? base + off : nullptr;
    off += (size_t)((bytes > 0 ? bytes : 16) + 255) & ~(size_t)255;
    return ptr;
  };
  b.G = (double*)take(chunk * n * n * 8);
  b.g0 = (double*)take(chunk * n * 8);
  b.CE = (double*)take(chunk * n * p * 8);
  b.ce0 = (double*)take(chunk * p * 8);
  b.CI = (double*)take(chunk * n * m * 8);
  b.ci0 = (double*)take(chunk * m * 8);
  b.x = (double*)take(chunk * n * 8);
  b.f = (double*)take(chunk * 8);
  b.status = (int*)take(chunk * 4);
  b.iq = (int*)take(chunk * 4);
  b.A = (int*)take(chunk * (p + m) * 4);
  b.iters = (int*)take(chunk * 4);
  b.CIv = (double*)take(chunk * m * 8);
  b.CIr = (unsigned char*)take(chunk * packed_row_stride(m));
  return off;
}

// One code byte per inequality column (mgqp_pack_unit_columns) -> the (value, row) pair the solvers read; in place
// on the row bytes.
__global__ void unit_code_expand_kernel(long long cnt, int m, int mr, unsigned char* __restrict__ row, double* __restrict__ val) {
  const long long total = cnt * m;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long q = e / m;
    const int i = (int)(e - q * m);
    const unsigned c = row[q * mr + i];
    const bool empty = (c & 0x40u) != 0;
    val[e] = empty ? 0.0 : ((c & 0x80u) ? -1.0 : 1.0);
    row[q * mr + i] = (unsigned char)(empty ? 0u : (c & 0x3fu));
  }
}

// Upload / solve / download QPs [q0, q0+cnt) of host batch h through buffers b on stream st.
// `pack`: the shape is served by the lane kernel and the caller did not opt out, so CI travels in
// column-compressed form whenever this chunk qualifies (every column a bound on one variable).
int run_chunk(const mgqp_batch* h, int device, DevBufs& b, long long q0, long long cnt, cudaStream_t st, bool pack,
              cudaEvent_t uploads_done, bool* try_unit) {
  const long long n = h->n, p = h->p, m = h->m;
  const long long mr = packed_row_stride(m);
  mgqp_batch d = *h;
  d.batch = cnt;
  d.CI_val = nullptr; d.CI_row = nullptr; d.sCIv = 0; d.sCIr = 0;
  bool ci_compressed = false;
  if (m > 0 && h->CI_val) {  // the caller's own compressed form: ship it as it is
    const long long cv = h->sCIv ? cnt : 1, cr = h->sCIr ? cnt : 1;
    if (h->sCIv == 0 || h->sCIv == m) {
      CU(cudaMemcpyAsync(b.CIv, h->CI_val + q0 * h->sCIv, (size_t)(cv * m) * 8, cudaMemcpyHostToDevice, st));
    } else {
      CU(cudaMemcpy2DAsync(b.CIv, (size_t)m * 8, h->CI_val + q0 * h->sCIv, (size_t)h->sCIv * 8, (size_t)m * 8, (size_t)cv,
                           cudaMemcpyHostToDevice, st));
    }
    CU(cudaMemcpy2DAsync(b.CIr, (size_t)mr, h->CI_row + q0 * h->sCIr, (size_t)(h->sCIr ? h->sCIr : m), (size_t)m, (size_t)cr,
                         cudaMemcpyHostToDevice, st));
    g_h2d_bytes += cv * m * 8 + cr * m;
    d.CI_val = b.CIv; d.CI_row = b.CIr;
    d.sCIv = h->sCIv ? m : 0; d.sCIr = h->sCIr ? mr : 0;
    ci_compressed = true;
  } else if (m > 0 && pack && b.hCIv && try_unit && *try_unit &&
             mgqp_pack_unit_columns((int32_t)n, (int32_t)m, cnt, h->CI + q0 * h->sCI, h->sCI, b.hCIr, mr, 0)) {
    // every coefficient is +-1: one byte per column over the link, expanded on the device
    CU(cudaMemcpyAsync(b.CIr, b.hCIr, (size_t)(cnt * mr), cudaMemcpyHostToDevice, st));
    g_h2d_bytes += cnt * mr;
    const long long total = cnt * m;
    const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    unit_code_expand_kernel<<<grid, 256, 0, st>>>(cnt, (int)m, (int)mr, b.CIr, b.CIv);
    g_launches.fetch_add(1);
    CU(cudaGetLastError());
    d.CI_val = b.CIv; d.CI_row = b.CIr; d.sCIv = m; d.sCIr = mr;
    ci_compressed = true;
  } else if (m > 0 && pack && b.hCIv && ((try_unit && *try_unit) ? (*try_unit = false, true) : true) &&
             mgqp_pack_bound_columns((int32_t)n, (int32_t)m, cnt, h->CI + q0 * h->sCI, h->sCI, b.hCIv, m, b.hCIr, mr, 0)) {
    CU(cudaMemcpyAsync(b.CIv, b.hCIv, (size_t)(cnt * m) * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(b.CIr, b.hCIr, (size_t)(cnt * mr), cudaMemcpyHostToDevice, st));
    g_h2d_bytes += cnt * m * 8 + cnt * mr;
    d.CI_val = b.CIv; d.CI_row = b.CIr; d.sCIv = m; d.sCIr = mr;
    ci_compressed = true;
  }
  struct F { const double* src; long long stride, size; double* dst; const double** dptr; int64_t* dstride; };
  F fs[6] = {{h->G, h->sG, n * n, b.G, &d.G, &d.sG},       {h->g0, h->sg0, n, b.g0, &d.g0, &d.sg0},
             {h->CE, h->sCE, n * p, b.CE, &d.CE, &d.sCE},  {h->ce0, h->sce0, p, b.ce0, &d.ce0, &d.sce0},
             {h->CI, h->sCI, n * m, b.CI, &d.CI, &d.sCI},  {h->ci0, h->sci0, m, b.ci0, &d.ci0, &d.sci0}};
  for (int k = 0; k < 6; k++) {
    F& f = fs[k];
    *f.dptr = f.dst;
    if (f.size == 0 || (k == 4 && ci_compressed)) { *f.dstride = 0; continue; }
    if (f.stride == 0) {  // broadcast field: one instance
      CU(cudaMemcpyAsync(f.dst, f.src, (size_t)f.size * 8, cudaMemcpyHostToDevice, st));
      *f.dstride = 0;
      g_h2d_bytes += f.size * 8;
    } else if (k < 2 && cnt > 1 && mgqp_blocks_all_equal(f.src + q0 * f.stride, f.stride, f.size, cnt, 0)) {
      // G / g0 identical in every QP of the chunk (the reference's caller: G = I, g0 = 0): one instance
      CU(cudaMemcpyAsync(f.dst, f.src + q0 * f.stride, (size_t)f.size * 8, cudaMemcpyHostToDevice, st));
      *f.dstride = 0;
      g_h2d_bytes += f.size * 8;
    } else if (f.stride == f.size) {
      CU(cudaMemcpyAsync(f.dst, f.src + q0 * f.stride, (size_t)(cnt * f.size) * 8, cudaMemcpyHostToDevice, st));
      *f.dstride = f.size;
      g_h2d_bytes += cnt * f.size * 8;
    } else {  // padded host layout: compact it on the way up
      CU(cudaMemcpy2DAsync(f.dst, (size_t)f.size * 8, f.src + q0 * f.stride, (size_t)f.stride * 8, (size_t)f.size * 8,
                           (size_t)cnt, cudaMemcpyHostToDevice, st));
      *f.dstride = f.size;
      g_h2d_bytes += cnt * f.size * 8;
    }
  }
  if (uploads_done) CU(cudaEventRecord(uploads_done, st));
  d.x = b.x; d.f = b.f; d.status = b.status;
  d.iq = h->iq ? b.iq : nullptr;
  d.A = h->A ? b.A : nullptr;
  d.iters = h->iters ? b.iters : nullptr;
  int rc = solve_device(&d, device, st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(h->x + q0 * n, b.x, (size_t)(cnt * n) * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h->f + q0, b.f, (size_t)cnt * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(h->status + q0, b.status, (size_t)cnt * 4, cudaMemcpyDeviceToHost, st));
  g_d2h_bytes += cnt * (n * 8 + 8 + 4);
  if (h->iq) { CU(cudaMemcpyAsync(h->iq + q0, b.iq, (size_t)cnt * 4, cudaMemcpyDeviceToHost, st)); g_d2h_bytes += cnt * 4; }
  if (h->A) {
    CU(cudaMemcpyAsync(h->A + q0 * (p + m), b.A, (size_t)(cnt * (p + m)) * 4, cudaMemcpyDeviceToHost, st));
    g_d2h_bytes += cnt * (p + m) * 4;
  }
  if (h->iters) { CU(cudaMemcpyAsync(h->iters + q0, b.iters, (size_t)cnt * 4, cudaMemcpyDeviceToHost, st)); g_d2h_bytes += cnt * 4; }
  return MGQP_OK;
}

long long default_chunk(const mgqp_batch* h) {
  const long long per_qp = mgqp_algorithmic_bytes(h->n, h->p, h->m);
  // 48k QPs per chunk, at most ~512 MB of inputs.  Measured (profiles/r02_notes.md 2c): the 7-variable shapes run
  // best at 32k-64k QPs per chunk (5.1e7 QP/s; 80k 4.9e7, 256k 4.4e7: shorter chunks fill the two-stage pipeline
  // sooner), the 14-variable ones want at least ~40k (S0: 8k 1.1e7, 21k 1.7e7, 42k 1.9e7 QP/s: per-chunk launch and
  // copy submission costs).
  long long c = (512ll << 20) / (per_qp > 0 ? per_qp : 1);
  if (c > 49152) c = 49152;
  if (c < 1024) c = 1024;
  c &= ~1023ll;
  return c;
}

// Pull chunks from `next` until the batch is exhausted; two streams, two buffer sets.
int pump(const mgqp_batch* h, int device, long long chunk, std::atomic<long long>* next, int ndev = 1,
         long long* chunks_done = nullptr) {
  int rc = validate(h);
  if (rc) return rc;
  DeviceCtx* ctx = nullptr;
  if ((rc = device_ctx(device, &ctx))) return rc;
  CU(cudaSetDevice(device));
  if (chunk > h->batch) chunk = h->batch;
  if (chunk < 1) return MGQP_OK;
  std::lock_guard<std::mutex> host_lock(ctx->host_mu);  // one host-buffer call per device at a time
  DevBufs probe, bufs[2];
  const size_t need = carve_bufs(probe, h, chunk, nullptr);
  for (int i = 0; i < 2; i++) {
    if (!ctx->host_st[i]) CU(cudaStreamCreateWithFlags(&ctx->host_st[i], cudaStreamNonBlocking));
    if (!ctx->host_up[i]) CU(cudaEventCreateWithFlags(&ctx->host_up[i], cudaEventDisableTiming));
    if (ctx->arena_cap[i] < need) {
      if (ctx->arena[i]) CU(cudaFree(ctx->arena[i]));
      ctx->arena[i] = nullptr;
      ctx->arena_cap[i] = 0;
      CU(cudaMalloc(&ctx->arena[i], need));
      ctx->arena_cap[i] = need;
    }
    carve_bufs(bufs[i], h, chunk, static_cast<char*>(ctx->arena[i]));
  }
  // CI goes up in column-compressed form where the lane kernel serves the shape (see run_chunk)
  bool pack = !(h->flags & (MGQP_FLAG_NO_HOST_PACK | MGQP_FLAG_CTA_KERNEL)) && !h->CI_val && h->m > 0 &&
              h->m <= 64 && h->sCI != 0 &&
              (make_lane_plan(h->n, h->p, h->m, ctx->smem_optin).ok || make_plan(h->n, h->p, h->m, ctx->smem_optin).gw > 0);
  if (pack) {
    // Worth it only if the host threads can read CI faster than the link would have carried it.  Measured
    // (profiles/r01_notes.md): ~6 GB/s of CI per packing thread, ~50 GB/s over PCIe; 16 threads per GPU
    // double the C2 rate, 8 still gain 28 %, 4 lose 20 % against dense transport.
    const double n = h->n, p = h->p, m = h->m;
    const double rest = 8.0 * (n * n + n + n * p + p + m) + 8.0 * (n + 1) + 4.0;
    const double link = 50e9, per_thread = 6e9;
    const double t_dense = (rest + 8.0 * n * m) / link;
    // (the devices of a multi-device call share the host's packing threads)
    const double threads = (double)mgqp_host_pack_threads() / (ndev > 0 ? ndev : 1);
    const double t_link = (rest + 9.0 * m) / link, t_cpu = 8.0 * n * m / (per_thread * (threads < 1.0 ? 1.0 : threads));
    pack = (t_link > t_cpu ? t_link : t_cpu) < 0.8 * t_dense;
  }
  if (pack) {
    const size_t vbytes = ((size_t)chunk * h->m * 8 + 255) & ~(size_t)255;
    const size_t pneed = vbytes + (size_t)chunk * packed_row_stride(h->m);
    for (int i = 0; i < 2; i++) {
      if (ctx->pack_cap[i] < pneed) {
        if (ctx->pack_host[i]) CU(cudaFreeHost(ctx->pack_host[i]));
        ctx->pack_host[i] = nullptr;
        ctx->pack_cap[i] = 0;
        CU(cudaHostAlloc(&ctx->pack_host[i], pneed, cudaHostAllocDefault));
        ctx->pack_cap[i] = pneed;
      }
      bufs[i].hCIv = static_cast<double*>(ctx->pack_host[i]);
      bufs[i].hCIr = static_cast<unsigned char*>(ctx->pack_host[i]) + vbytes;
    }
  }
  cudaStream_t* st = ctx->host_st;
  // the one-byte form of +-1 columns is tried until a chunk of this call turns out not to have it
  bool try_unit = pack && h->n <= 64 && mgqp_pack_unit_is_fast() && !std::getenv("MGQP_NO_UNIT_CODES");
  int k = 0;
  while (!rc) {
    const long long q0 = next->fetch_add(chunk);
    if (q0 >= h->batch) break;
    const long long cnt = (h->batch - q0) < chunk ? (h->batch - q0) : chunk;
    // The buffer set is reused every second chunk.  Its device buffers are protected by stream order; what the host
    // has to wait for is only the pinned staging area the packer is about to overwrite, i.e. the UPLOADS of the
    // set's previous chunk -- not that chunk's kernel and downloads, which a stream synchronisation would add to
    // the packer's critical path (the call was bound by exactly that path: profiles/r02_notes.md).
    static const bool full_sync = std::getenv("MGQP_HOST_STREAM_SYNC") != nullptr;  // A/B knob: the former behaviour
    if (full_sync && cudaStreamSynchronize(st[k & 1]) != cudaSuccess) { rc = fail(MGQP_ERR_CUDA, "stream sync failed"); break; }
    if (k >= 2 && cudaEventSynchronize(ctx->host_up[k & 1]) != cudaSuccess) { rc = fail(MGQP_ERR_CUDA, "event sync failed"); break; }
    rc = run_chunk(h, device, bufs[k & 1], q0, cnt, st[k & 1], pack, ctx->host_up[k & 1], &try_unit);
    k++;
    if (chunks_done) ++*chunks_done;
  }
  for (int i = 0; i < 2; i++) {
    cudaError_t e = cudaStreamSynchronize(st[i]);
    if (e != cudaSuccess && !rc) rc = fail(MGQP_ERR_CUDA, "stream sync failed: %s", cudaGetErrorString(e));
  }
  return rc;
}

// ------------------------------------------------------------------ the solver's callers, batched
int cascade_device(const mgqp_scenarios* s, int device, cudaStream_t st) {
  if (!s) return fail(MGQP_ERR_INVALID_ARG, "null scenario descriptor");
  if (s->batch < 0) return fail(MGQP_ERR_INVALID_ARG, "batch = %lld must be >= 0", (long long)s->batch);
  if (!s->q || !s->qd || !s->M || !s->h || !s->J || !s->Jd || !s->pos_des || !s->pos_cur || !s->vel_des ||
      !s->vel_cur || !s->acc_des || !s->acc_cur || !s->q1_des || !s->tracking || !s->torques)
    return fail(MGQP_ERR_INVALID_ARG, "null scenario field");
  if (s->joint_task < 0 || s->joint_task >= kDof) return fail(MGQP_ERR_INVALID_ARG, "joint_task out of range");
  if (s->batch == 0) return MGQP_OK;
  DeviceCtx* ctx = nullptr;
  int rc = device_ctx(device, &ctx);
  if (rc) return rc;
  CU(cudaSetDevice(device));
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->casc_const) {
      // [G = I14 (196) | g0 = 0 (14) | CI_val (28) | CI_row (28 bytes, padded)]
      std::vector<double> hc(196 + 14 + 28 + 4, 0.0);
      for (int i = 0; i < kN; i++) hc[i * kN + i] = 1.0;
      for (int c = 0; c < kN; c++) { hc[210 + c] = -1.0; hc[210 + kN + c] = 1.0; }   // src/mgqp.cpp:1111-1112
      unsigned char* rows = reinterpret_cast<unsigned char*>(hc.data() + 238);
      for (int c = 0; c < kM; c++) rows[c] = (unsigned char)(c % kN);
      CU(cudaMalloc((void**)&ctx->casc_const, hc.size() * 8));
      CU(cudaMemcpy(ctx->casc_const, hc.data(), hc.size() * 8, cudaMemcpyHostToDevice));
    }
  }
  const double* cG = ctx->casc_const;
  const double* cg0 = cG + 196;
  const double* cCIv = cG + 210;
  const uint8_t* cCIr = reinterpret_cast<const uint8_t*>(cG + 238);
  // workspace per scenario: doubles CE0 140 | ce00 10 | ci0 28 | x0 14 | f0 1 | CE2 14 | ce02 1 | CI2 392 | x2 14 | f2 1,
  // floats goal2 1 | Z 196 | y 14, ints st0 | st2
  const long long perD = 140 + 10 + 28 + 14 + 1 + 14 + 1 + 392 + 14 + 1, perF = 1 + 196 + 14, perI = 2;
  const long long chunk_max = 1ll << 17;
  for (long long q0 = 0; q0 < s->batch; q0 += chunk_max) {
    const long long B = (s->batch - q0) < chunk_max ? (s->batch - q0) : chunk_max;
    AsyncBuf ws;
    ws.st = st;
    CU(cudaMallocFromPoolAsync(&ws.p, (size_t)B * (perD * 8 + perF * 4 + perI * 4) + 256, ctx->pool, st));
    double* d = static_cast<double*>(ws.p);
    CascadeParams P;
    memset(&P, 0, sizeof P);
    P.batch = B;
    P.q = s->q + q0 * 7; P.qd = s->qd + q0 * 7; P.M = s->M + q0 * 49; P.h = s->h + q0 * 7;
    P.J = s->J + q0 * 21; P.Jd = s->Jd + q0 * 21;
    P.pos_des = s->pos_des + q0 * 3; P.pos_cur = s->pos_cur + q0 * 3; P.vel_des = s->vel_des + q0 * 3;
    P.vel_cur = s->vel_cur + q0 * 3; P.acc_des = s->acc_des + q0 * 3; P.acc_cur = s->acc_cur + q0 * 3;
    P.q1_des = s->q1_des + q0;
    P.kp_task = s->kp_task; P.kd_task = s->kd_task; P.kp_joint = s->kp_joint; P.kd_joint = s->kd_joint;
    for (int i = 0; i < kDof; i++) {
      P.acc_p[i] = s->acc_p[i]; P.acc_n[i] = s->acc_n[i]; P.tau_p[i] = s->tau_p[i]; P.tau_n[i] = s->tau_n[i];
      P.q_sup[i] = s->q_sup[i]; P.q_inf[i] = s->q_inf[i];
    }
    P.joint_task = s->joint_task;
    double* x0 = nullptr; double* f0 = nullptr; double* x2 = nullptr; double* f2 = nullptr;
    P.CE0 = d; d += B * 140;
    P.ce00 = d; d += B * 10;
    P.ci0 = d; d += B * 28;
    x0 = d; d += B * 14;
    f0 = d; d += B;
    P.CE2 = d; d += B * 14;
    P.ce02 = d; d += B;
    P.CI2 = d; d += B * 392;
    x2 = d; d += B * 14;
    f2 = d; d += B;
    float* fl = reinterpret_cast<float*>(d);
    P.goal2 = fl; fl += B;
    P.Z = fl; fl += B * 196;
    P.y = fl; fl += B * 14;
    int* st0 = reinterpret_cast<int*>(fl);
    int* st2 = st0 + B;
    P.x0 = x0; P.f0 = f0; P.st0 = st0; P.x2 = x2; P.f2 = f2; P.st2 = st2;
    P.tracking = s->tracking + q0 * kN; P.torques = s->torques + q0 * kDof;
    P.status0 = s->status0 ? s->status0 + q0 : nullptr;
    P.status2 = s->status2 ? s->status2 + q0 : nullptr;
    const unsigned blocks = (unsigned)((B + 127) / 128);
    cascade_assemble_kernel<<<blocks, 128, 0, st>>>(P);
    g_launches.fetch_add(1);
    mgqp_batch l0;
    memset(&l0, 0, sizeof l0);
    l0.n = kN; l0.p = kP0; l0.m = kM; l0.batch = B;
    l0.G = cG; l0.g0 = cg0; l0.CE = P.CE0; l0.ce0 = P.ce00; l0.CI = nullptr; l0.ci0 = P.ci0;
    l0.sG = 0; l0.sg0 = 0; l0.sCE = 140; l0.sce0 = 10; l0.sCI = 0; l0.sci0 = 28;
    l0.CI_val = cCIv; l0.CI_row = cCIr; l0.sCIv = 0; l0.sCIr = 0;
    l0.x = x0; l0.f = f0; l0.status = st0;
    l0.flags = MGQP_FLAG_RETRY_WITHOUT_CI;
    if ((rc = solve_device(&l0, device, st))) return rc;
    // (64 threads: 64 x 140 floats of working matrices stay under the 48 KB a launch may ask for unannounced)
    cascade_project_kernel<<<(unsigned)((B + 63) / 64), 64, 64 * kN * kP0 * sizeof(float), st>>>(P);
    g_launches.fetch_add(1);
    {
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return fail(MGQP_ERR_CUDA, "cascade_project_kernel launch failed: %s", cudaGetErrorString(e));
    }
    mgqp_batch l2;
    memset(&l2, 0, sizeof l2);
    l2.n = kN; l2.p = 1; l2.m = kM; l2.batch = B;
    l2.G = cG; l2.g0 = cg0; l2.CE = P.CE2; l2.ce0 = P.ce02; l2.CI = P.CI2; l2.ci0 = P.ci0;
    l2.sG = 0; l2.sg0 = 0; l2.sCE = 14; l2.sce0 = 1; l2.sCI = 392; l2.sci0 = 28;
    l2.x = x2; l2.f = f2; l2.status = st2;
    l2.flags = MGQP_FLAG_RETRY_WITHOUT_CI | kFlagDenseCI;  // limitsMatrix * Z: dense rows of rank 4
    if ((rc = solve_device(&l2, device, st))) return rc;
    cascade_finish_kernel<<<blocks, 128, 0, st>>>(P);
    g_launches.fetch_add(1);
    if (s->tap_CE0) CU(cudaMemcpyAsync(s->tap_CE0 + q0 * 140, P.CE0, (size_t)B * 140 * 8, cudaMemcpyDeviceToDevice, st));
    if (s->tap_ce0) CU(cudaMemcpyAsync(s->tap_ce0 + q0 * 10, P.ce00, (size_t)B * 10 * 8, cudaMemcpyDeviceToDevice, st));
    if (s->tap_ci0) CU(cudaMemcpyAsync(s->tap_ci0 + q0 * 28, P.ci0, (size_t)B * 28 * 8, cudaMemcpyDeviceToDevice, st));
    if (s->tap_Z) CU(cudaMemcpyAsync(s->tap_Z + q0 * 196, P.Z, (size_t)B * 196 * 4, cudaMemcpyDeviceToDevice, st));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(MGQP_ERR_CUDA, "cascade kernels failed: %s", cudaGetErrorString(e));
  }
  return MGQP_OK;
}

// host pointers: upload the state of a chunk, run the cascade, download the answers; two streams alternate
int cascade_host(const mgqp_scenarios* h, int device, long long chunk) {
  if (!h || h->batch < 0) return fail(MGQP_ERR_INVALID_ARG, "bad scenario descriptor");
  if (h->batch == 0) return MGQP_OK;
  DeviceCtx* ctx = nullptr;
  int rc = device_ctx(device, &ctx);
  if (rc) return rc;
  CU(cudaSetDevice(device));
  if (chunk <= 0) chunk = 1 << 16;
  if (chunk > h->batch) chunk = h->batch;
  std::lock_guard<std::mutex> host_lock(ctx->host_mu);
  const int inF[13] = {7, 7, 49, 7, 21, 21, 3, 3, 3, 3, 3, 3, 1};
  const float* const src[13] = {h->q, h->qd, h->M, h->h, h->J, h->Jd, h->pos_des, h->pos_cur, h->vel_des,
                                h->vel_cur, h->acc_des, h->acc_cur, h->q1_des};
  long long in_floats = 0;
  for (int k = 0; k < 13; k++) in_floats += inF[k];
  const size_t need = (size_t)chunk * (in_floats * 4 + 14 * 4 + 7 * 4 + 8) + 16 * 256;
  for (int i = 0; i < 2; i++) {
    if (!ctx->casc_st[i]) CU(cudaStreamCreateWithFlags(&ctx->casc_st[i], cudaStreamNonBlocking));
    if (ctx->casc_cap[i] < need) {
      if (ctx->casc_arena[i]) CU(cudaFree(ctx->casc_arena[i]));
      ctx->casc_arena[i] = nullptr; ctx->casc_cap[i] = 0;
      CU(cudaMalloc(&ctx->casc_arena[i], need));
      ctx->casc_cap[i] = need;
    }
  }
  g_h2d_bytes = g_d2h_bytes = 0;
  int k = 0;
  for (long long q0 = 0; q0 < h->batch && !rc; q0 += chunk, k++) {
    const long long B = (h->batch - q0) < chunk ? (h->batch - q0) : chunk;
    cudaStream_t st = ctx->casc_st[k & 1];
    CU(cudaStreamSynchronize(st));
    char* base = static_cast<char*>(ctx->casc_arena[k & 1]);
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base + off; off += (bytes + 255) & ~(size_t)255; return p; };
    mgqp_scenarios d = *h;
    d.batch = B;
    const float* dev_in[13];
    for (int f = 0; f < 13; f++) {
      float* p = reinterpret_cast<float*>(take((size_t)B * inF[f] * 4));
      CU(cudaMemcpyAsync(p, src[f] + q0 * inF[f], (size_t)B * inF[f] * 4, cudaMemcpyHostToDevice, st));
      dev_in[f] = p;
      g_h2d_bytes += B * inF[f] * 4;
    }
    d.q = dev_in[0]; d.qd = dev_in[1]; d.M = dev_in[2]; d.h = dev_in[3]; d.J = dev_in[4]; d.Jd = dev_in[5];
    d.pos_des = dev_in[6]; d.pos_cur = dev_in[7]; d.vel_des = dev_in[8]; d.vel_cur = dev_in[9];
    d.acc_des = dev_in[10]; d.acc_cur = dev_in[11]; d.q1_des = dev_in[12];
    d.tracking = reinterpret_cast<float*>(take((size_t)B * 14 * 4));
    d.torques = reinterpret_cast<float*>(take((size_t)B * 7 * 4));
    d.status0 = h->status0 ? reinterpret_cast<int32_t*>(take((size_t)B * 4)) : nullptr;
    d.status2 = h->status2 ? reinterpret_cast<int32_t*>(take((size_t)B * 4)) : nullptr;
    d.tap_CE0 = nullptr; d.tap_ce0 = nullptr; d.tap_ci0 = nullptr; d.tap_Z = nullptr;
    if ((rc = cascade_device(&d, device, st))) break;
    CU(cudaMemcpyAsync(h->tracking + q0 * 14, d.tracking, (size_t)B * 14 * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->torques + q0 * 7, d.torques, (size_t)B * 7 * 4, cudaMemcpyDeviceToHost, st));
    g_d2h_bytes += B * 21 * 4;
    if (h->status0) { CU(cudaMemcpyAsync(h->status0 + q0, d.status0, (size_t)B * 4, cudaMemcpyDeviceToHost, st)); g_d2h_bytes += B * 4; }
    if (h->status2) { CU(cudaMemcpyAsync(h->status2 + q0, d.status2, (size_t)B * 4, cudaMemcpyDeviceToHost, st)); g_d2h_bytes += B * 4; }
  }
  for (int i = 0; i < 2; i++) {
    cudaError_t e = cudaStreamSynchronize(ctx->casc_st[i]);
    if (e != cudaSuccess && !rc) rc = fail(MGQP_ERR_CUDA, "stream sync failed: %s", cudaGetErrorString(e));
  }
  return rc;
}

}  // namespace

// ====================================================================== exported C ABI
extern "C" {

int solve_quadprog_batched(const mgqp_batch* d, int device, void* stream) {
  DeviceGuard guard;
  return solve_device(d, device, static_cast<cudaStream_t>(stream));
}

int solve_quadprog_batched_host(const mgqp_batch* h, int device, int64_t chunk) {
  g_h2d_bytes = g_d2h_bytes = 0;
  int rc = validate(h);
  if (rc) return rc;
  if (chunk <= 0) chunk = default_chunk(h);
  std::atomic<long long> next(0);
  DeviceGuard guard;
  return pump(h, device, chunk, &next);
}

int solve_quadprog_batched_multi(const mgqp_batch* h, const int* devices, int ndev, int64_t chunk) {
  int rc = validate(h);
  if (rc) return rc;
  if (!devices || ndev < 1) return fail(MGQP_ERR_INVALID_ARG, "need at least one device");
  if (chunk <= 0) {
    chunk = default_chunk(h);
    // at least ~4 chunks per device so the dynamic sharding can balance
    while (chunk > 1024 && h->batch / chunk < 4ll * ndev) chunk /= 2;
  }
  std::atomic<long long> next(0);
  std::atomic<long long> up(0), down(0);
  std::vector<int> rcs(ndev, 0);
  std::vector<std::string> errs(ndev);
  std::vector<long long> chunks(ndev, 0);
  std::vector<std::thread> th;
  for (int i = 0; i < ndev; i++)
    th.emplace_back([&, i]() {
      g_h2d_bytes = g_d2h_bytes = 0;
      rcs[i] = pump(h, devices[i], chunk, &next, ndev, &chunks[i]);
      if (rcs[i]) errs[i] = g_last_error;
      up += g_h2d_bytes;
      down += g_d2h_bytes;
    });
  for (auto& t : th) t.join();
  g_multi_chunks = chunks;
  g_h2d_bytes = up.load();
  g_d2h_bytes = down.load();
  for (int i = 0; i < ndev; i++)
    if (rcs[i]) return fail(rcs[i], "device %d: %s", devices[i], errs[i].c_str());
  return MGQP_OK;
}

int mgqp_cascade_batched(const mgqp_scenarios* s, int device, void* stream) {
  DeviceGuard guard;
  return cascade_device(s, device, static_cast<cudaStream_t>(stream));
}

int mgqp_cascade_batched_host(const mgqp_scenarios* s, int device, int64_t chunk) {
  DeviceGuard guard;
  return cascade_host(s, device, chunk);
}

int mgqp_solve_one(int32_t n, int32_t p, int32_t m, const double* G, const double* g0, const double* CE,
                   const double* ce0, const double* CI, const double* ci0, double* x, double* f, int32_t* status,
                   double* G_factor, int device) {
  // One stream, one pinned and one device staging block per device, kept across calls: a call is one H2D
  // copy, one kernel launch (static schedule: no work-queue memset), one D2H copy and one synchronisation.
  if (n < 1 || p < 0 || m < 0 || !G || !g0 || !x || !f || !status || (p && (!CE || !ce0)) || (m && (!CI || !ci0)))
    return fail(MGQP_ERR_INVALID_ARG, "mgqp_solve_one: bad argument");
  DeviceCtx* ctx = nullptr;
  int rc = device_ctx(device, &ctx);
  if (rc) return rc;
  DeviceGuard guard;
  CU(cudaSetDevice(device));
  const size_t sz[6] = {(size_t)n * n, (size_t)n, (size_t)n * p, (size_t)p, (size_t)n * m, (size_t)m};
  const double* src[6] = {G, g0, CE, ce0, CI, ci0};
  size_t off[7];
  off[0] = 0;
  for (int k = 0; k < 6; k++) off[k + 1] = off[k] + ((sz[k] + 1) & ~(size_t)1);  // 16-byte aligned fields
  const size_t in_dbl = off[6], out_dbl = (size_t)n + 2;                          // x | f | status
  std::lock_guard<std::mutex> lk(ctx->one_mu);
  if (!ctx->one_st) CU(cudaStreamCreateWithFlags(&ctx->one_st, cudaStreamNonBlocking));
  if (ctx->one_cap < in_dbl + out_dbl) {
    if (ctx->one_host) cudaFreeHost(ctx->one_host);
    if (ctx->one_dev) cudaFree(ctx->one_dev);
    ctx->one_host = nullptr; ctx->one_dev = nullptr; ctx->one_cap = 0;
    const size_t cap = (in_dbl + out_dbl) * 2 + 64;
    CU(cudaHostAlloc((void**)&ctx->one_host, cap * 8, cudaHostAllocDefault));
    CU(cudaMalloc((void**)&ctx->one_dev, cap * 8));
    ctx->one_cap = cap;
  }
  double* hb = ctx->one_host;
  double* db = ctx->one_dev;
  for (int k = 0; k < 6; k++)
    if (sz[k]) memcpy(hb + off[k], src[k], sz[k] * 8);
  cudaStream_t st = ctx->one_st;
  CU(cudaMemcpyAsync(db, hb, in_dbl * 8, cudaMemcpyHostToDevice, st));
  mgqp_batch d;
  memset(&d, 0, sizeof d);
  d.n = n; d.p = p; d.m = m; d.batch = 1;
  d.G = db + off[0]; d.g0 = db + off[1]; d.CE = db + off[2]; d.ce0 = db + off[3]; d.CI = db + off[4]; d.ci0 = db + off[5];
  d.sG = (int64_t)sz[0]; d.sg0 = n; d.sCE = (int64_t)sz[2]; d.sce0 = p; d.sCI = (int64_t)sz[4]; d.sci0 = m;
  d.x = db + in_dbl; d.f = db + in_dbl + n; d.status = reinterpret_cast<int32_t*>(db + in_dbl + n + 1);
  d.flags = MGQP_FLAG_STATIC_SCHEDULE;
  if ((rc = solve_device(&d, device, st))) return rc;
  CU(cudaMemcpyAsync(hb + in_dbl, db + in_dbl, out_dbl * 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  memcpy(x, hb + in_dbl, (size_t)n * 8);
  *f = hb[in_dbl + n];
  memcpy(status, hb + in_dbl + n + 1, 4);
  g_h2d_bytes = (long long)in_dbl * 8;
  g_d2h_bytes = (long long)out_dbl * 8;
  if (G_factor) {
    // The reference leaves chol(G) in G (lower factor mirrored into the upper triangle,
    // QuadProg++.hh:42-45).  n is tiny on this path; the factor is recomputed on the host in the
    // reference's operation order so the caller sees the same bits.
    for (int i = 0; i < n * n; i++) G_factor[i] = G[i];
    for (int i = 0; i < n; i++) {
      for (int j = i; j < n; j++) {
        double sum = G_factor[i * n + j];
        for (int k = i - 1; k >= 0; k--) sum -= G_factor[i * n + k] * G_factor[j * n + k];
        if (i == j) {
          if (sum <= 0.0) return MGQP_OK;  // status already says NOT_PD
          G_factor[i * n + i] = sqrt(sum);
        } else {
          G_factor[j * n + i] = sum / G_factor[i * n + i];
        }
      }
      for (int k = i + 1; k < n; k++) G_factor[i * n + k] = G_factor[k * n + i];
    }
  }
  return MGQP_OK;
}

const char* mgqp_status_string(int status) {
  switch (status) {
    case MGQP_STATUS_OK: return "ok";
    case MGQP_STATUS_INFEASIBLE: return "infeasible";
    case MGQP_STATUS_NOT_PD: return "G not positive definite";
    case MGQP_STATUS_EQ_DEPENDENT: return "Constraints are linearly dependent";
    case MGQP_STATUS_ITER_CAP: return "iteration cap reached";
    default: return "unknown";
  }
}

const char* mgqp_last_error(void) { return g_last_error.c_str(); }

void mgqp_last_transfer_bytes(int64_t* h2d, int64_t* d2h) {
  if (h2d) *h2d = g_h2d_bytes;
  if (d2h) *d2h = g_d2h_bytes;
}

int mgqp_last_multi_chunks(int64_t* counts, int max_devices) {
  const int n = (int)g_multi_chunks.size();
  for (int i = 0; i < n && i < max_devices; i++) counts[i] = g_multi_chunks[i];
  return n;
}

namespace mgqp {
__global__ void div_check_kernel(long long count, unsigned long long seed, unsigned long long* out) {
  unsigned long long acc = 0, bad = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    auto mix = [](unsigned long long z) {
      z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
      z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
      return z ^ (z >> 31);
    };
    const unsigned long long ra = mix(seed + 2 * (unsigned long long)i * 0x9e3779b97f4a7c15ull);
    const unsigned long long rb = mix(seed + (2 * (unsigned long long)i + 1) * 0x9e3779b97f4a7c15ull);
    double a, b;
    const int mode = (int)(i % 3);
    if (mode == 0) {  // raw bit patterns: every exponent, denormals, inf, NaN
      a = __longlong_as_double((long long)ra);
      b = __longlong_as_double((long long)rb);
    } else {          // exponents within 2^+-40 of one (mode 1) / of each other anywhere in the range (mode 2)
      const unsigned long long ea = 1023ull - 40 + (ra >> 12) % 81, eb = 1023ull - 40 + (rb >> 12) % 81;
      const unsigned long long sh = (mode == 2) ? (mix(ra ^ rb) % 1800) : 0;
      const unsigned long long xa = (mode == 2) ? (ea + sh > 983 ? ea + sh - 983 : 1) : ea;
      const unsigned long long xb = (mode == 2) ? (eb + sh > 983 ? eb + sh - 983 : 1) : eb;
      a = __longlong_as_double((long long)((ra & 0x800fffffffffffffull) | ((xa & 0x7ff) << 52)));
      b = __longlong_as_double((long long)((rb & 0x800fffffffffffffull) | ((xb & 0x7ff) << 52)));
    }
    bool ok = true;
    const double r = d_rcp_refined(b);
    const double q = d_div_rcp(a, b, r, ok);
    const double a2 = __longlong_as_double((long long)mix(ra));  // a second numerator over the same refinement
    bool ok2 = true;
    const double q2 = d_div_rcp(a2, b, r, ok2);
    if (ok) {
      acc++;
      if (__double_as_longlong(q) != __double_as_longlong(__ddiv_rn(a, b))) bad++;
    }
    if (ok2) {
      acc++;
      if (__double_as_longlong(q2) != __double_as_longlong(__ddiv_rn(a2, b))) bad++;
    }
    bool ok3 = true;
    const double rt = d_sqrt_split(a, ok3);
    if (ok3) {
      acc++;
      if (__double_as_longlong(rt) != __double_as_longlong(__dsqrt_rn(a))) bad++;
    }
  }
  atomicAdd(out, acc);
  atomicAdd(out + 1, bad);
}
}  // namespace mgqp

int mgqp_debug_div_check(int64_t count, uint64_t seed, int64_t* accepted, int64_t* mismatches) {
  unsigned long long* d = nullptr;
  unsigned long long h[2] = {0, 0};
  cudaError_t e = cudaMalloc(&d, sizeof(h));
  if (e == cudaSuccess) e = cudaMemset(d, 0, sizeof(h));
  if (e == cudaSuccess) {
    mgqp::div_check_kernel<<<1184, 256>>>((long long)count, (unsigned long long)seed, d);
    g_launches.fetch_add(1);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  if (d) cudaFree(d);
  if (e != cudaSuccess) {
    g_last_error = std::string("mgqp_debug_div_check: ") + cudaGetErrorString(e);
    return MGQP_ERR_CUDA;
  }
  if (accepted) *accepted = (int64_t)h[0];
  if (mismatches) *mismatches = (int64_t)h[1];
  return MGQP_OK;
}

int mgqp_device_count(void) {
  int c = 0;
  if (cudaGetDeviceCount(&c) != cudaSuccess) return 0;
  return c;
}

int64_t mgqp_kernel_launch_count(void) { return g_launches.load(); }

int mgqp_plan(int32_t n, int32_t p, int32_t m, int32_t* smem_bytes_per_cta, int32_t* warps_per_cta) {
  const LanePlan lp = make_lane_plan(n, p, m, kSmemBudget);
  if (lp.ok) {
    if (smem_bytes_per_cta) *smem_bytes_per_cta = lp.smem;
    if (warps_per_cta) *warps_per_cta = lp.warps;
    return 1;
  }
  const Plan pl = make_plan(n, p, m, kSmemBudget);
  if (pl.gw < 0) {
    const LargePlan lg = make_large_plan(n, p, m, kSmemBudget);
    if (!lg.ok) return -2;
    if (smem_bytes_per_cta) *smem_bytes_per_cta = lg.v.smem_bytes;
    if (warps_per_cta) *warps_per_cta = lg.T / 32;
    return 0;
  }
  if (smem_bytes_per_cta) *smem_bytes_per_cta = pl.smem;
  if (warps_per_cta) *warps_per_cta = pl.warps;
  return pl.gw;
}

int64_t mgqp_algorithmic_bytes(int32_t n, int32_t p, int32_t m) {
  const int64_t N = n, P = p, M = m;
  return 8 * (N * N + N + N * P + P + N * M + M) + 8 * (N + 1) + 4;
}

}  // extern "C"
