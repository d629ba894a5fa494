// TEST INFRASTRUCTURE: CPU stand-in for csrc/warp_prims.cuh.
//
// The 32 lanes of one warp run as ucontext fibers inside one OS thread.  Every warp primitive
// (shuffle, ballot, any, reduce, syncwarp) is a rendezvous: the lane publishes its operand into a
// double-buffered slot, yields to the scheduler, and reads the peers' operands when it is resumed
// after every lane has published.  The scheduler also checks that all 32 lanes arrived at the SAME
// primitive (kind + sequence number) -- a lane that took a different path through a sync point is
// reported instead of silently mis-synchronising, which is exactly the class of bug a lock-step
// warp kernel can have.  Arithmetic is plain IEEE fp64 (build with -ffp-contract=off), i.e. the
// same one-rounding-per-operation semantics as the __d*_rn intrinsics of the device build.
#pragma once
#ifndef MGQP_EMULATE
#error "warp_emu.h is for the -DMGQP_EMULATE host build only"
#endif

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#define MGQP_DEV static inline
#define MGQP_MEMBER inline
#define MGQP_DEV_NOINLINE static inline
#define MGQP_UNROLL1
#define MGQP_UNROLL2
#ifndef MGQP_UNROLL4
#define MGQP_UNROLL4
#endif
#define MGQP_HD
#define MGQP_HOST
#define MGQP_FULL 0xffffffffu

namespace mgqp {

#define MGQP_EMU_MAX 128
struct EmuWarp {
  ucontext_t sched;
  ucontext_t ctx[MGQP_EMU_MAX];
  char* stack[MGQP_EMU_MAX];
  bool finished[MGQP_EMU_MAX];
  uint64_t slot[2][MGQP_EMU_MAX];
  uint32_t kind[2][MGQP_EMU_MAX];
  uint32_t seq[MGQP_EMU_MAX];
  int cur;
  int nthreads;
  void (*body)(void*);
  void* arg;
  long rendezvous;
};

static thread_local EmuWarp* g_emu = nullptr;

static void emu_trampoline() {
  EmuWarp* w = g_emu;
  w->body(w->arg);
  w->finished[w->cur] = true;
  swapcontext(&w->ctx[w->cur], &w->sched);
}

// run body(arg) on 32 lock-step lanes
static inline void emu_run_block(void (*body)(void*), void* arg, int nthreads) {
  static thread_local EmuWarp* w = nullptr;
  const size_t STK = 256 * 1024;
  if (!w) {
    w = (EmuWarp*)calloc(1, sizeof(EmuWarp));
    for (int i = 0; i < MGQP_EMU_MAX; i++) w->stack[i] = (char*)malloc(STK);
  }
  g_emu = w;
  w->body = body;
  w->arg = arg;
  w->rendezvous = 0;
  w->nthreads = nthreads;
  const int NT = nthreads;
  for (int i = 0; i < NT; i++) {
    w->finished[i] = false;
    w->seq[i] = 0;
    getcontext(&w->ctx[i]);
    w->ctx[i].uc_stack.ss_sp = w->stack[i];
    w->ctx[i].uc_stack.ss_size = STK;
    w->ctx[i].uc_link = &w->sched;
    makecontext(&w->ctx[i], (void (*)())emu_trampoline, 0);
  }
  for (;;) {
    int alive = 0;
    for (int i = 0; i < NT; i++) {
      if (w->finished[i]) continue;
      alive++;
      w->cur = i;
      swapcontext(&w->sched, &w->ctx[i]);
    }
    if (!alive) break;
    // after a full round every unfinished lane sits at a rendezvous: check they agree
    int first = -1;
    for (int i = 0; i < NT; i++) {
      if (w->finished[i]) continue;
      if (first < 0) first = i;
      const int par = (w->seq[i] - 1) & 1;
      const int par0 = (w->seq[first] - 1) & 1;
      if (w->seq[i] != w->seq[first] || w->kind[par][i] != w->kind[par0][first]) {
        fprintf(stderr, "EMU: lanes %d and %d diverged at a warp primitive (seq %u/%u kind %u/%u)\n", first, i,
                w->seq[first], w->seq[i], w->kind[par0][first], w->kind[par][i]);
        abort();
      }
    }
    int fin = 0;
    for (int i = 0; i < NT; i++) fin += w->finished[i];
    if (fin != 0 && fin != NT) {
      fprintf(stderr, "EMU: %d lanes exited while others wait at a warp primitive\n", fin);
      abort();
    }
    w->rendezvous++;
  }
}

static inline void emu_run_warp(void (*body)(void*), void* arg) { emu_run_block(body, arg, 32); }

// publish + yield; returns the parity buffer holding this round's operands
static inline int emu_meet(uint32_t kind, uint64_t v) {
  EmuWarp* w = g_emu;
  const int me = w->cur;
  const int par = w->seq[me] & 1;
  w->slot[par][me] = v;
  w->kind[par][me] = kind;
  w->seq[me]++;
  swapcontext(&w->ctx[me], &w->sched);
  w->cur = me;
  return par;
}

static inline uint64_t emu_bits(double v) { uint64_t b; memcpy(&b, &v, 8); return b; }
static inline double emu_dbl(uint64_t b) { double v; memcpy(&v, &b, 8); return v; }

MGQP_DEV int wp_lane() { return g_emu->cur; }
// thread-block view of the same fibers (gi_large.cuh)
MGQP_DEV int blk_tid() { return g_emu->cur; }
MGQP_DEV int blk_dim() { return g_emu->nthreads; }
MGQP_DEV void blk_sync() { emu_meet(20, 0); }
MGQP_DEV void wp_sync() { emu_meet(1, 0); }
MGQP_DEV bool wp_any(bool p) {
  const int par = emu_meet(2, p);
  bool r = false;
  for (int i = 0; i < 32; i++) r = r || g_emu->slot[par][i];
  return r;
}
MGQP_DEV unsigned wp_ballot(bool p) {
  const int par = emu_meet(3, p);
  unsigned r = 0;
  for (int i = 0; i < 32; i++) r |= (g_emu->slot[par][i] ? 1u : 0u) << i;
  return r;
}
MGQP_DEV int wp_max(int v) {
  const int par = emu_meet(4, (uint64_t)(int64_t)v);
  int r = INT32_MIN;
  for (int i = 0; i < 32; i++) { int o = (int)(int64_t)g_emu->slot[par][i]; if (o > r) r = o; }
  return r;
}
MGQP_DEV int wp_min(int v) {
  const int par = emu_meet(5, (uint64_t)(int64_t)v);
  int r = INT32_MAX;
  for (int i = 0; i < 32; i++) { int o = (int)(int64_t)g_emu->slot[par][i]; if (o < r) r = o; }
  return r;
}
static inline uint64_t emu_shfl_idx(uint32_t kind, uint64_t v, int src_abs_or_self) {
  const int par = emu_meet(kind, v);
  return g_emu->slot[par][src_abs_or_self];
}
static inline int emu_src(int src, int width) {
  const int me = g_emu->cur;
  return (me & ~(width - 1)) + (src & (width - 1));
}
static inline int emu_src_down(int d, int width) {
  const int me = g_emu->cur;
  const int li = me & (width - 1);
  return (li + d < width) ? me + d : me;
}
static inline int emu_src_xor(int mask, int width) {
  const int me = g_emu->cur;
  const int li = (me & (width - 1)) ^ mask;
  return (li < width) ? (me & ~(width - 1)) + li : me;
}
MGQP_DEV double wp_shfl(double v, int src, int width) { return emu_dbl(emu_shfl_idx(6, emu_bits(v), emu_src(src, width))); }
MGQP_DEV int wp_shfl(int v, int src, int width) { return (int)(int64_t)emu_shfl_idx(7, (uint64_t)(int64_t)v, emu_src(src, width)); }
MGQP_DEV double wp_shfl_down(double v, int d, int width) { return emu_dbl(emu_shfl_idx(8, emu_bits(v), emu_src_down(d, width))); }
MGQP_DEV int wp_shfl_down(int v, int d, int width) { return (int)(int64_t)emu_shfl_idx(9, (uint64_t)(int64_t)v, emu_src_down(d, width)); }
MGQP_DEV double wp_shfl_xor(double v, int mask, int width) { return emu_dbl(emu_shfl_idx(10, emu_bits(v), emu_src_xor(mask, width))); }
MGQP_DEV int wp_shfl_xor(int v, int mask, int width) { return (int)(int64_t)emu_shfl_idx(11, (uint64_t)(int64_t)v, emu_src_xor(mask, width)); }

MGQP_DEV double d_mul(double a, double b) { return a * b; }
MGQP_DEV double d_add(double a, double b) { return a + b; }
MGQP_DEV double d_sub(double a, double b) { return a - b; }
MGQP_DEV double d_div(double a, double b) { return a / b; }
MGQP_DEV double d_div_z(double a, double b) { return a / b; }
// the split inline path of the device's division (warp_prims.cuh): here the quotient is the host's IEEE division and
// `ok` follows the device's range tests, so that the callers' fallbacks are exercised on the CPU as well
MGQP_DEV double d_sqrt_split(double x, bool& ok) {
  unsigned long long u;
  memcpy(&u, &x, 8);
  ok = ok && ((unsigned)((unsigned)(u >> 32) - 0x03500000u) < 0x7ca00000u);
  return sqrt(x);
}
MGQP_DEV double d_rcp_refined(double) { return 0.0; }
MGQP_DEV double d_div_rcp(double a, double b, double, bool& ok) {
  const double q = a / b;
  auto hi = [](double v) { unsigned long long u; memcpy(&u, &v, 8); return (unsigned)(u >> 32) & 0x7fffffffu; };
  ok = ok && hi(q) > 0x00100000u && hi(q) <= 0x7f800000u && hi(b) < 0x7f800000u && hi(a) >= 0x03600000u &&
       hi(a) <= 0x7f800000u;
  return q;
}
MGQP_DEV double d_sqrt(double a) { return sqrt(a); }
MGQP_DEV double d_fma(double a, double b, double c) { return fma(a, b, c); }
MGQP_DEV bool d_signbit(double a) { return __builtin_signbit(a) != 0; }
MGQP_DEV double d_abs(double a) { return fabs(a); }
MGQP_DEV void d_acc_neg(double& acc, double a) { if (a < 0.0) acc = acc + a; }
MGQP_DEV void d_or_if_nonzero(unsigned& mask, double a, unsigned bit) { if (a != 0.0) mask |= bit; }
MGQP_DEV int i_popc(unsigned v) { return __builtin_popcount(v); }
MGQP_DEV double d_inf() { return INFINITY; }
MGQP_DEV bool d_is_pos_zero(double a) { return emu_bits(a) == 0; }
MGQP_DEV int wp_ffs(unsigned v) { return __builtin_ffs((int)v); }

}  // namespace mgqp
