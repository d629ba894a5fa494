// TEST INFRASTRUCTURE: runs the DEVICE solver body (csrc/gi_small.cuh) on the CPU under the fiber
// warp emulator, so that its lock-step logic can be compared bit-for-bit with the oracle here,
// where there is no GPU.  Build: g++ -O1 -ffp-contract=off -DMGQP_EMULATE -shared -fPIC ...
#include "warp_emu.h"
#include "../../motion-generation-using-quadratic-programs_b200/csrc/gi_small.cuh"

#include <thread>
#include <vector>
#include <atomic>

using namespace mgqp;

namespace {
struct Job {
  TileView tv;
  double* ws;
  int cnt, max_iter, gw, fma;
  TileOut out;
};

void lane_body(void* a) {
  Job* j = (Job*)a;
  if (j->fma) {
    if (j->gw == 8) gi_solve_tile<8, true>(j->tv, j->ws, j->cnt, j->max_iter, j->out);
    else if (j->gw == 16) gi_solve_tile<16, true>(j->tv, j->ws, j->cnt, j->max_iter, j->out);
    else gi_solve_tile<32, true>(j->tv, j->ws, j->cnt, j->max_iter, j->out);
  } else {
    if (j->gw == 8) gi_solve_tile<8, false>(j->tv, j->ws, j->cnt, j->max_iter, j->out);
    else if (j->gw == 16) gi_solve_tile<16, false>(j->tv, j->ws, j->cnt, j->max_iter, j->out);
    else gi_solve_tile<32, false>(j->tv, j->ws, j->cnt, j->max_iter, j->out);
  }
}
}  // namespace

extern "C" int emu_solve_batch(int n, int p, int m, long long batch, const double* G, const double* g0,
                               const double* CE, const double* ce0, const double* CI, const double* ci0,
                               double* x, double* f, int* status, int* iq, int* A, int* iters, int max_iter,
                               int gw, int fma, int threads) {
  if (n > gw || m > 32 * gw) return -2;
  const TileView tv = make_tile_view(n, p, m, gw);
  const int qpw = tv.qpw;
  const long long tiles = (batch + qpw - 1) / qpw;
  if (max_iter <= 0) max_iter = 100 + 10 * (p + m);
  std::atomic<long long> next(0);
  auto work = [&]() {
    std::vector<double> ws(tv.total);
    for (;;) {
      const long long t = next.fetch_add(1);
      if (t >= tiles) break;
      const long long q0 = t * qpw;
      const int cnt = (int)((batch - q0) < qpw ? (batch - q0) : qpw);
      for (int g = 0; g < qpw; g++) {
        const long long q = q0 + (g < cnt ? g : 0);  // tail groups replicate the tile's first QP
        memcpy(&ws[tv.oG + g * n * n], G + q * n * n, sizeof(double) * n * n);
        memcpy(&ws[tv.og0 + g * n], g0 + q * n, sizeof(double) * n);
        memcpy(&ws[tv.oCE + g * n * p], CE + q * n * p, sizeof(double) * n * p);
        memcpy(&ws[tv.oce0 + g * p], ce0 + q * p, sizeof(double) * p);
        memcpy(&ws[tv.oCI + g * n * m], CI + q * n * m, sizeof(double) * n * m);
        memcpy(&ws[tv.oci0 + g * m], ci0 + q * m, sizeof(double) * m);
      }
      Job j;
      j.tv = tv; j.ws = ws.data(); j.cnt = cnt; j.max_iter = max_iter; j.gw = gw; j.fma = fma;
      j.out.x = x + q0 * n; j.out.f = f + q0; j.out.status = status + q0;
      j.out.iq = iq ? iq + q0 : nullptr;
      j.out.A = A ? A + q0 * (p + m) : nullptr;
      j.out.iters = iters ? iters + q0 : nullptr;
      emu_run_warp(lane_body, &j);
    }
  };
  if (threads <= 1) { work(); return 0; }
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) pool.emplace_back(work);
  for (auto& t : pool) t.join();
  return 0;
}

// The same with the inequality block kept column-compressed in the (emulated) shared memory, as the kernel keeps a
// batch that arrives as mgqp_batch::CI_val / CI_row: CIv [batch][m] values, CIr [batch][m] rows.
extern "C" int emu_solve_batch_comp(int n, int p, int m, long long batch, const double* G, const double* g0,
                                    const double* CE, const double* ce0, const double* CIv, const unsigned char* CIr,
                                    const double* ci0, double* x, double* f, int* status, int* iq, int* A, int* iters,
                                    int max_iter, int gw, int fma) {
  if (n > gw || m > 32 * gw) return -2;
  const TileView tv = make_tile_view(n, p, m, gw, 1);
  const int qpw = tv.qpw;
  const long long tiles = (batch + qpw - 1) / qpw;
  if (max_iter <= 0) max_iter = 100 + 10 * (p + m);
  std::vector<double> ws(tv.total);
  for (long long t = 0; t < tiles; t++) {
    const long long q0 = t * qpw;
    const int cnt = (int)((batch - q0) < qpw ? (batch - q0) : qpw);
    for (int g = 0; g < qpw; g++) {
      const long long q = q0 + (g < cnt ? g : 0);
      memcpy(&ws[tv.oG + g * n * n], G + q * n * n, sizeof(double) * n * n);
      memcpy(&ws[tv.og0 + g * n], g0 + q * n, sizeof(double) * n);
      memcpy(&ws[tv.oCE + g * n * p], CE + q * n * p, sizeof(double) * n * p);
      memcpy(&ws[tv.oce0 + g * p], ce0 + q * p, sizeof(double) * p);
      double* ci = &ws[tv.oCI + g * tv.cisz];
      for (int i = 0; i < m; i++) {
        ci[i] = CIv[q * m + i];
        unsigned r = CIr[q * m + i];
        reinterpret_cast<int*>(ci + m)[i] = (int)(r < (unsigned)n ? r : (unsigned)(n - 1));
      }
      memcpy(&ws[tv.oci0 + g * m], ci0 + q * m, sizeof(double) * m);
    }
    Job j;
    j.tv = tv; j.ws = ws.data(); j.cnt = cnt; j.max_iter = max_iter; j.gw = gw; j.fma = fma;
    j.out.x = x + q0 * n; j.out.f = f + q0; j.out.status = status + q0;
    j.out.iq = iq ? iq + q0 : nullptr;
    j.out.A = A ? A + q0 * (p + m) : nullptr;
    j.out.iters = iters ? iters + q0 : nullptr;
    emu_run_warp(lane_body, &j);
  }
  return 0;
}

// ------------------------------------------------------------------ lane-per-QP solver (gi_lane.cuh)
#include "../../motion-generation-using-quadratic-programs_b200/csrc/gi_lane.cuh"

namespace {
struct LaneJob {
  LaneView v;
  double* ws;
  LaneTileIO tio;
  int max_iter, fma;
  double* x; double* f; int* status; int* iq; int* A; int* iters;  // of the tile's first QP
};

template <int N>
void lane_tile_n(LaneJob* j) {
  const int lane = wp_lane();
  double xo[8];
  double fo;
  int st, q, it;
  if (j->fma) gi_lane_tile<N, true>(j->v, j->ws, j->tio, j->max_iter, xo, fo, st, q, it);
  else gi_lane_tile<N, false>(j->v, j->ws, j->tio, j->max_iter, xo, fo, st, q, it);
  if (lane >= j->tio.cnt) return;
  const int n = j->v.n, p = j->v.p, m = j->v.m;
  for (int i = 0; i < n; i++) j->x[lane * n + i] = xo[i];
  j->f[lane] = fo;
  j->status[lane] = st;
  if (j->iq) j->iq[lane] = q;
  if (j->iters) j->iters[lane] = it;
  const signed char* Ab = reinterpret_cast<const signed char*>(j->ws + lane * j->v.stride + j->v.oByte) + j->v.bA;
  if (j->A) for (int i = 0; i < p + m; i++) j->A[lane * (p + m) + i] = (i < q) ? (int)Ab[i] : 0x7fffffff;
}

void lane_tile_body(void* a) {
  LaneJob* j = (LaneJob*)a;
  switch (j->v.n) {
    case 1: lane_tile_n<1>(j); break;
    case 2: lane_tile_n<2>(j); break;
    case 3: lane_tile_n<3>(j); break;
    case 4: lane_tile_n<4>(j); break;
    case 5: lane_tile_n<5>(j); break;
    case 6: lane_tile_n<6>(j); break;
    case 7: lane_tile_n<7>(j); break;
    default: lane_tile_n<8>(j); break;
  }
}
}  // namespace

// CIv / CIr (optional): the column-compressed form of CI (mgqp_batch::CI_val / CI_row), [batch][m] each
extern "C" int emu_lane_solve_batch_c(int n, int p, int m, long long batch, const double* G, const double* g0,
                                      const double* CE, const double* ce0, const double* CI, const double* ci0,
                                      double* x, double* f, int* status, int* iq, int* A, int* iters, int max_iter,
                                      int fma, const double* CIv, const unsigned char* CIr);

extern "C" int emu_lane_solve_batch(int n, int p, int m, long long batch, const double* G, const double* g0,
                                    const double* CE, const double* ce0, const double* CI, const double* ci0,
                                    double* x, double* f, int* status, int* iq, int* A, int* iters, int max_iter,
                                    int fma) {
  return emu_lane_solve_batch_c(n, p, m, batch, G, g0, CE, ce0, CI, ci0, x, f, status, iq, A, iters, max_iter, fma,
                                nullptr, nullptr);
}

// The warp's 32 lanes run as fibers: the cooperative stage-in phases of gi_lane_tile synchronise the warp.
extern "C" int emu_lane_solve_batch_c(int n, int p, int m, long long batch, const double* G, const double* g0,
                                      const double* CE, const double* ce0, const double* CI, const double* ci0,
                                      double* x, double* f, int* status, int* iq, int* A, int* iters, int max_iter,
                                      int fma, const double* CIv, const unsigned char* CIr) {
  if (n < 1 || n > 8 || m > 64 || p > n) return -2;
  const LaneView v = make_lane_view(n, p, m);
  if (max_iter <= 0) max_iter = 100 + 10 * (p + m);
  const long long tiles = (batch + 31) / 32;
  std::atomic<long long> next(0);
  auto work = [&]() {
    std::vector<double> ws((size_t)32 * v.stride);
    for (;;) {
      const long long t = next.fetch_add(1);
      if (t >= tiles) break;
      const long long q0 = t * 32;
      for (size_t i = 0; i < ws.size(); i++) ws[i] = -1e300;  // poison
      LaneJob j;
      j.v = v; j.ws = ws.data(); j.max_iter = max_iter; j.fma = fma;
      const double* in[6] = {G + q0 * n * n, g0 + q0 * n, CE + q0 * n * p, ce0 + q0 * p,
                             CI ? CI + q0 * (long long)n * m : nullptr, ci0 + q0 * m};
      const long long st[6] = {(long long)n * n, n, (long long)n * p, p, (long long)n * m, m};
      for (int k = 0; k < 6; k++) { j.tio.in[k] = in[k]; j.tio.stride[k] = st[k]; }
      if (CIv) { j.tio.CIv = CIv + q0 * m; j.tio.CIr = CIr + q0 * m; j.tio.sCIv = m; j.tio.sCIr = m; }
      j.tio.cnt = (int)((batch - q0) < 32 ? (batch - q0) : 32);
      j.x = x + q0 * n; j.f = f + q0; j.status = status + q0;
      j.iq = iq ? iq + q0 : nullptr;
      j.A = A ? A + q0 * (p + m) : nullptr;
      j.iters = iters ? iters + q0 : nullptr;
      emu_run_warp(lane_tile_body, &j);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 0; t < 4; t++) pool.emplace_back(work);
  for (auto& t : pool) t.join();
  return 0;
}

// ------------------------------------------------------------------ CTA-per-QP solver (gi_large.cuh)
#include "../../motion-generation-using-quadratic-programs_b200/csrc/gi_large.cuh"

namespace {
struct LargeJob {
  LargeView v;
  unsigned char* smem;
  double* ws;
  LaneIO io;
  int max_iter, fma;
  LargeOut out;
};
void large_body(void* a) {
  LargeJob* j = (LargeJob*)a;
  if (j->fma) gi_large_solve<true>(j->v, j->smem, j->ws, j->io, j->max_iter, j->out);
  else gi_large_solve<false>(j->v, j->smem, j->ws, j->io, j->max_iter, j->out);
}
}  // namespace

extern "C" int emu_large_solve_batch(int n, int p, int m, long long batch, const double* G, const double* g0,
                                     const double* CE, const double* ce0, const double* CI, const double* ci0,
                                     double* x, double* f, int* status, int* iq, int* A, int* iters, int max_iter,
                                     int nthreads, int fma, int threads) {
  if (nthreads > MGQP_EMU_MAX || (nthreads & (nthreads - 1))) return -2;
  const LargeView v = make_large_view(n, p, m, nthreads);
  if (max_iter <= 0) max_iter = 100 + 10 * (p + m);
  std::atomic<long long> next(0);
  auto work = [&]() {
    std::vector<unsigned char> smem(v.smem_bytes + 16);
    std::vector<double> ws(v.ws_doubles);
    for (;;) {
      const long long b = next.fetch_add(1);
      if (b >= batch) break;
      LargeJob j;
      j.v = v;
      j.smem = (unsigned char*)(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
      j.ws = ws.data();
      j.io = LaneIO{G + b * n * n, g0 + b * n, CE + b * n * p, ce0 + b * p, CI + b * (long long)n * m, ci0 + b * m};
      j.max_iter = max_iter;
      j.fma = fma;
      j.out.x = x + b * n; j.out.f = f + b; j.out.status = status + b;
      j.out.iq = iq ? iq + b : nullptr;
      j.out.A = A ? A + b * (p + m) : nullptr;
      j.out.iters = iters ? iters + b : nullptr;
      emu_run_block(large_body, &j, nthreads);
    }
  };
  if (threads <= 1) { work(); return 0; }
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) pool.emplace_back(work);
  for (auto& t : pool) t.join();
  return 0;
}
