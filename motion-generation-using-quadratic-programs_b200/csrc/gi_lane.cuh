// Goldfarb-Idnani dual active-set solver, ONE QP PER LANE (n <= 8, m <= 64): a warp advances 32
// independent QPs in SIMT fashion.
//
// Replaces the body of the reference's solve_quadprog (lib/QuadProgpp/libquadprog.a: QuadProg++.o,
// prototype include/QuadProgpp/QuadProg++.hh:69-72, behavioural spec SURVEY.md Appendix A) for the
// small joint-space shapes (BASELINE.json configs[1]/[4]: n=7, p=1, m=42).
//
// Why a second kernel family: with n = 7 a lane group of 8 (gi_small.cuh) spends most of its issue
// slots on work that is uniform across the group (divisions, square roots, control) and on ordered
// cross-lane sums; one lane per QP executes each of those once per QP.
//
// A tile (32 QPs) goes through five phases:
//   A  cooperative, coalesced stage-in of G and g0: consecutive lanes copy consecutive doubles of the
//      tile's block (cp.async, 8 bytes each) into the OWNER lane's private shared-memory slot;
//   S  lane-owned set-up (Cholesky of G, J = L^-T, x0 = -G^-1 g0) in the reference's operation order;
//   B  cooperative, coalesced classification of the inequality block: lane l takes column c of QP q
//      ((q, c) enumerate the tile's 32*m columns), reads its n coefficients -- 32 lanes read 32
//      consecutive doubles of a row of CI -- and writes ONE byte into the owner's slot:
//        * a single nonzero coefficient equal to +1 or -1 -> the index of +x_k / -x_k in a small table
//          (every inequality the reference itself ever builds: src/mgqp.cpp:1111-1112,
//          limitsMatrix = [-I; +I]);
//        * no nonzero coefficient                          -> the index of the table's 0.0 entry;
//        * anything else (GEN)                              -> the column is re-read from global memory
//          (L2) whenever it is scanned, with the reference's full n-term sum.
//      ci0 is copied into the slot by cp.async on the way;
//   I  the active-set iteration, lane-owned: J lives in REGISTERS, the rest of a lane's state in its slot;
//   O  results (x through shared memory so that the global stores are coalesced).
// With the +-1 columns reduced to one byte each the slot is 107 doubles at n = 7, m = 42 (it was 149
// when every column kept a coefficient), which is what lets eight warps share an SM.
//
// The byte form is used ONLY in the violated-constraint scan, and only while x is finite:
// sum_j CI[j][i]*x[j] with exact-zero coefficients and a +-1 equals +-x[k] bit for bit (up to the sign
// of a zero, which no comparison can see).  Everything else (compute_d, step lengths, the s[ip]
// refresh) works on the dense column np = CI[:,ip].
//
// Each lane runs the reference's scalar recurrences in the reference's own operation order, so
// results are bit-identical to QuadProg++ by construction (no cross-lane arithmetic at all).
//
// Slot layout (doubles; the lane stride is odd so the 16 lanes of a half warp hit 16 distinct
// 8-byte bank pairs whenever they address the same element):
//   area A: R (upper Hessenberg, packed) | d[N] | r[N] | zero | u[N+1]      (set-up: J scratch | g0/y/x)
//   area B: ci0[m] | int8: A[N+1], A_old[N+1], gen, cidx[m]                 (set-up: L)
// d | r | zero double as the scan's table [x | -x | 0.0] (d and r are dead while a scan runs).
#pragma once
#include "warp_prims.cuh"

#ifndef MGQP_EPS
#define MGQP_EPS 2.220446049250313e-16
#endif

#ifdef MGQP_EMULATE
#define MGQP_LDG(p) (*(p))
#define MGQP_LDG2(p, a, b) do { (a) = (p)[0]; (b) = (p)[1]; } while (0)
#define MGQP_LDG2_STREAM(p, a, b) MGQP_LDG2(p, a, b)
#define MGQP_PREFETCH_L2(p) do { } while (0)
#define MGQP_PREFETCH_L2_KEEP(p) do { } while (0)
#define MGQP_CP8(dst, src) (*(dst) = *(src))
#define MGQP_CP8_AT(sbase, base, off, src) ((base)[off] = *(src))
#define MGQP_SHARED_BASE(p) 0u
#define MGQP_CP_WAIT() do { } while (0)
#define MGQP_OPAQUE_PTR(p) do { } while (0)
#else
// hide a pointer's derivation from the optimiser, so that `p + i` stays ONE 64-bit multiply-add on a 32-bit i
// instead of being re-associated into 64-bit index arithmetic
#define MGQP_OPAQUE_PTR(p) asm volatile("" : "+l"(p))
#define MGQP_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#define MGQP_PREFETCH_L2_KEEP(p) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p))
#define MGQP_LDG(p) __ldg(p)
#define MGQP_LDG2(p, a, b) do { const double2 v2_ = __ldg(reinterpret_cast<const double2*>(p)); (a) = v2_.x; (b) = v2_.y; } while (0)
#define MGQP_LDG2_STREAM(p, a, b) asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p))
// 8-byte asynchronous copy global -> shared: no register, no scoreboard; completion by MGQP_CP_WAIT
#define MGQP_CP8(dst, src) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory")
// the same with the 32-bit shared-window address of `base` taken once outside the loop (the conversion costs four
// uniform-datapath instructions each time it is rematerialised)
#define MGQP_SHARED_BASE(p) ((unsigned)__cvta_generic_to_shared(p))
#define MGQP_CP8_AT(sbase, base, off, src) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((sbase) + 8u * (unsigned)(off)), "l"(src) : "memory")
#define MGQP_CP_WAIT() asm volatile("cp.async.wait_all;" ::: "memory")
#endif

#ifdef MGQP_LANE_KEEP
#define MGQP_PREFETCH_NEXT(p) MGQP_PREFETCH_L2_KEEP(p)
#else
#define MGQP_PREFETCH_NEXT(p) MGQP_PREFETCH_L2(p)
#endif
#ifndef MGQP_LANE_K
#define MGQP_LANE_K 1
#endif
#ifndef MGQP_LANE_SWEEP_GROUP
#define MGQP_LANE_SWEEP_GROUP 3
#endif
// 1 = the scan reads a second byte per column that already points at +inf for columns it may not choose (two
// instructions fewer per column).  Measured: -1.2 % on C2, -3.3 % on C5 -- the slot grows from 107 to 113 doubles
// (231 KB of shared memory per CTA, 13 KB less L1) and every flag change pays a refresh.  Off.
#ifndef MGQP_LANE_EFF_BYTE
#define MGQP_LANE_EFF_BYTE 0
#endif
// 1 = ci0 is loaded with the column's rows and stored, instead of a cp.async per column (which ptxas pads with
// three no-ops).  Measured: 7.485 against 7.430 ms.  Off.
#ifndef MGQP_LANE_CI0_LDG
#define MGQP_LANE_CI0_LDG 0
#endif
#ifndef MGQP_LANE_GEN_AFTER
#define MGQP_LANE_GEN_AFTER 1
#endif
#ifndef MGQP_LANE_PREFETCH_NEXT_CI
#define MGQP_LANE_PREFETCH_NEXT_CI 0  // measured: 1 = -3 %, 2 = -9 % (the lines are evicted before phase B wants them)
#endif
#ifndef MGQP_LANE_SQRT_SPLIT
#define MGQP_LANE_SQRT_SPLIT 1
#endif
#ifndef MGQP_LANE_GIVENS_CHAIN
#define MGQP_LANE_GIVENS_CHAIN 0
#endif
#ifndef MGQP_LANE_GIVENS_SPLIT
#define MGQP_LANE_GIVENS_SPLIT 1
#endif
#ifndef MGQP_LANE_SPLIT_SWEEP
#define MGQP_LANE_SPLIT_SWEEP 0
#endif
#ifndef MGQP_LANE_UNROLLED_SETUP
#define MGQP_LANE_UNROLLED_SETUP 1
#endif

namespace mgqp {

struct LaneView {
  int n, p, m;
  int oR, od, orr, ozero, ou, oci0, oByte;            // doubles
  int oJs, ov, oLs;                                    // set-up scratch (doubles)
  int bA, bAold, bgen, bcidx, beff;                    // bytes, relative to (signed char*)(slot + oByte)
  int oinf;                                            // double: +inf (MGQP_LANE_EFF_BYTE)
  int stride;                                          // doubles per lane, odd
};

// R is kept in upper-Hessenberg packed form: row i stores columns >= i-1 (delete_constraint shifts
// columns left, which puts one entry below the diagonal until the Givens sweep removes it).
MGQP_HD inline int lane_rh_size(int n) { return (n - 1) * n - ((n - 1) * (n - 2)) / 2 + n; }
#define MGQP_RH(i, j) ((i) * N - (((i) * ((i)-1)) >> 1) + (j))

MGQP_HOST inline LaneView make_lane_view(int n, int p, int m) {
  LaneView v;
  v.n = n; v.p = p; v.m = m;
  int o = 0;
  v.oR = o;     o += lane_rh_size(n);
  v.od = o;     o += n;
  v.orr = o;    o += n;
  v.ozero = o;  o += 1;
  v.ou = o;     o += n + 1;
  // set-up scratch in area A: J (n*n) | g0 -> y -> x (n)
  v.oJs = 0;
  v.ov = n * n;
  if (o < n * n + n) o = n * n + n;
  v.oci0 = o;
  v.oLs = o;    // set-up scratch in area B: the Cholesky factor (n*n)
#if MGQP_LANE_EFF_BYTE
  const int bytes = 2 * (n + 1) + 1 + 2 * m;
#else
  const int bytes = 2 * (n + 1) + 1 + m;
#endif
  int b = m + (bytes + 7) / 8;
  if (b < n * n) b = n * n;
  v.oByte = o + m;
  o += b;
  v.bA = 0;
  v.bAold = n + 1;
  v.bgen = 2 * (n + 1);
  v.bcidx = 2 * (n + 1) + 1;
  v.beff = v.bcidx + m;
  v.oinf = o;
#if MGQP_LANE_EFF_BYTE
  o += 1;      // +inf: the entry of the scan's table that columns in the active set / excluded columns point at
#endif
  if ((o & 1) == 0) o++;  // odd stride: the 16 lanes of a half warp land on 16 distinct 8-byte bank pairs
  v.stride = o;
  return v;
}

// per-column byte: bits 3-7 = byte offset (index * 8) into the table [x_0..x_{N-1} | -x_0..-x_{N-1} | 0.0]
// of the term the column contributes to s_i (GEN columns: the row of the single nonzero, where known);
// GEN = the column needs the reference's full sum; EXCL (iaexcl == false) | ACT (iai == -1).
// A column may be chosen iff (byte & (EXCL | ACT)) == 0.
enum { LN_GEN = 0x01, LN_EXCL = 0x02, LN_ACT = 0x04, LN_IDX = 0xf8, LN_IDX_SHIFT = 3 };

#if MGQP_LANE_SQRT_SPLIT
#define LN_SQRT(v) d_sqrt_split(v, ok)
#else
#define LN_SQRT(v) d_sqrt(v)
#endif

MGQP_DEV double ln_distance(double a, double b) {
  // One division and one square root whichever argument is the larger: the two branches of the reference
  // are the same formula with the arguments swapped, and lanes that disagree on the comparison would
  // otherwise run both.  Equal magnitudes -- and the reference's fall-through for unordered ones -- return
  // a1*sqrt(2) up front: sparse problems (G = I, bound constraints) meet distance(0, 0) all the time.
  const double a1 = d_abs(a), b1 = d_abs(b);
  const bool agt = a1 > b1;
  if (!agt && !(b1 > a1)) return d_mul(a1, d_sqrt(2.0));
  const double hi = agt ? a1 : b1, lo = agt ? b1 : a1;
  const double t = d_div(lo, hi);  // (the zero-numerator shortcut d_div_z costs this kernel 1.3 %: its d is dense)
  return d_mul(hi, d_sqrt(d_add(1.0, d_mul(t, t))));
}

MGQP_DEV_NOINLINE double ln_distance_ni(double a, double b) { return ln_distance(a, b); }

// The same value on the split inline paths of the division and the square root (warp_prims.cuh): no convergence
// region on the way, which is what a serial chain of hypotenuses pays for (the CTA-per-QP kernel's add_constraint
// runs ~n of them on one thread); outside the inline paths' range the plain form takes over.
MGQP_DEV double ln_distance_split(double a, double b) {
  const double a1 = d_abs(a), b1 = d_abs(b);
  const bool agt = a1 > b1;
  if (!agt && !(b1 > a1)) return d_mul(a1, d_sqrt(2.0));
  const double hi = agt ? a1 : b1, lo = agt ? b1 : a1;
  bool ok = true;
  const double t = d_div_rcp(lo, hi, d_rcp_refined(hi), ok);
  const double h = d_mul(hi, d_sqrt_split(d_add(1.0, d_mul(t, t)), ok));
  if (ok) return h;
  return ln_distance(a, b);
}

// (c, s) -> normalised rotation coefficients exactly as add_constraint / delete_constraint derive
// them (Appendix A).  One shared, non-inlined copy: it holds four IEEE divisions and a square root,
// and the unrolled sweeps call it from twelve places.  valid == 0 <=> the reference `continue`s.
struct LaneGivens {
  double c, s, xny, diag;
  int valid;
};
MGQP_DEV_NOINLINE LaneGivens ln_givens(double cc, double ss) {
  LaneGivens g;
#if MGQP_LANE_GIVENS_SPLIT
  // The four divisions on the split inline path of the division (warp_prims.cuh): no branch between them, c and s
  // share the refinement of 1 / h, and one test at the end sends the whole call to the form below when any of the
  // quotients needs the division's general routine (exact zeros in d, operands at the ends of the exponent range).
  {
    bool ok = true;
    const double a1 = d_abs(cc), b1 = d_abs(ss);
    const bool agt = a1 > b1;
    const bool differ = agt || (b1 > a1);
    const double hi = differ ? (agt ? a1 : b1) : 1.0, lo = differ ? (agt ? b1 : a1) : 1.0;
    const double t = d_div_rcp(lo, hi, d_rcp_refined(hi), ok);
    const double h = differ ? d_mul(hi, LN_SQRT(d_add(1.0, d_mul(t, t)))) : d_mul(a1, d_sqrt(2.0));
    const double rh = d_rcp_refined(h);
    double s1 = d_div_rcp(ss, h, rh, ok);
    double c1 = d_div_rcp(cc, h, rh, ok);
    double dg = h;
    if (c1 < 0.0) {
      c1 = -c1;
      s1 = -s1;
      dg = -h;
    }
    const double den = d_add(1.0, c1);
    const double xq = d_div_rcp(d_abs(s1), den, d_rcp_refined(den), ok);  // s / den == -((-s) / den) exactly
    if (ok) {
      g.valid = !(d_abs(h) < MGQP_EPS);
      g.c = c1;
      g.s = s1;
      g.diag = dg;
      g.xny = d_signbit(s1) ? -xq : xq;
      return g;
    }
  }
#endif
  const double h = ln_distance(cc, ss);
  g.valid = !(d_abs(h) < MGQP_EPS);
  ss = d_div(ss, h);
  cc = d_div(cc, h);
  if (cc < 0.0) {
    cc = -cc;
    ss = -ss;
    g.diag = -h;
  } else {
    g.diag = h;
  }
  g.c = cc;
  g.s = ss;
  g.xny = d_div(ss, d_add(1.0, cc));
  return g;
}

// The add_constraint sweep chains its rotations through one number only: rotation j-1 starts from
// distance(d[j-2], |diag_j|), and |diag_j| is rotation j's hypotenuse.  This call does rotation j's coefficients
// (given its hypotenuse h) AND rotation j-1's hypotenuse, two independent dependency chains of about equal length
// in one block of straight-line code, which halves the latency of the sweep.  Same operations on the same operands
// as ln_givens / ln_distance; the split division's single range test sends the call to their plain forms.
struct LaneGivensNext {
  double c, s, xny, diag, hnext;
  int valid;
};
MGQP_DEV_NOINLINE LaneGivensNext ln_givens_next(double cc, double ss, double h, double anext) {
  LaneGivensNext g;
  g.valid = !(d_abs(h) < MGQP_EPS);
  const double bnext = g.valid ? h : cc;  // an invalid rotation leaves d[j-1] what it was
  {
    bool ok = true;
    // rotation j-1's hypotenuse
    const double a1 = d_abs(anext), b1 = d_abs(bnext);
    const bool agt = a1 > b1;
    const bool differ = agt || (b1 > a1);
    const double hi = differ ? (agt ? a1 : b1) : 1.0, lo = differ ? (agt ? b1 : a1) : 1.0;
    const double t = d_div_rcp(lo, hi, d_rcp_refined(hi), ok);
    const double hn = differ ? d_mul(hi, LN_SQRT(d_add(1.0, d_mul(t, t)))) : d_mul(a1, d_sqrt(2.0));
    // rotation j's coefficients
    const double rh = d_rcp_refined(h);
    double s1 = d_div_rcp(ss, h, rh, ok);
    double c1 = d_div_rcp(cc, h, rh, ok);
    double dg = h;
    if (c1 < 0.0) {
      c1 = -c1;
      s1 = -s1;
      dg = -h;
    }
    const double den = d_add(1.0, c1);
    const double xq = d_div_rcp(d_abs(s1), den, d_rcp_refined(den), ok);  // s / den == -((-s) / den) exactly
    if (ok) {
      g.c = c1;
      g.s = s1;
      g.diag = dg;
      g.xny = d_signbit(s1) ? -xq : xq;
      g.hnext = hn;
      return g;
    }
  }
  g.hnext = ln_distance(anext, bnext);
  ss = d_div(ss, h);
  cc = d_div(cc, h);
  if (cc < 0.0) {
    cc = -cc;
    ss = -ss;
    g.diag = -h;
  } else {
    g.diag = h;
  }
  g.c = cc;
  g.s = ss;
  g.xny = d_div(ss, d_add(1.0, cc));
  return g;
}

struct LaneIO {  // global memory, this lane's QP
  const double* G; const double* g0; const double* CE; const double* ce0; const double* CI; const double* ci0;
  // lane kernel only: column-compressed CI (mgqp_batch::CI_val / CI_row); CI is unused when CIv is set
  const double* CIv = nullptr;
  const unsigned char* CIr = nullptr;
};

// Global memory of one tile: the fields of the tile's FIRST QP and the element strides between QPs.
struct LaneTileIO {
  const double* in[6];        // G, g0, CE, ce0, CI, ci0
  long long stride[6];
  const double* CIv = nullptr;  // column-compressed CI (mgqp_batch::CI_val / CI_row) or null
  const unsigned char* CIr = nullptr;
  long long sCIv = 0, sCIr = 0;
  int cnt = 32;               // QPs in the tile (lanes >= cnt replicate QP 0 and store nothing)
  const double* next_G = nullptr;   // the G / g0 blocks of the tile this warp will take next (L2 prefetch), or null
  const double* next_g0 = nullptr;
  const double* next_CI = nullptr;  // ... and its CI / ci0 blocks
  const double* next_ci0 = nullptr;
};

// Flat enumeration of (QP q, element k) pairs of a tile with L elements per QP: lane l takes pairs
// l, l + 32, ...; running offsets into the global block (element stride sg between QPs) and into the
// slots (stride ss), no division or 64-bit multiply inside the loops.
struct LaneWalk {
  int q, k, L, dq, dk;
  long long soff, sstep, swrap;
  int doff, dstep, dwrap;
  MGQP_MEMBER void init(int lane, int L_, long long sg, int ss) {
    L = L_;
    dq = 32 / L; dk = 32 - dq * L;
    q = lane / L; k = lane - q * L;
    soff = q * sg + k; sstep = dq * sg + dk; swrap = sg - L;
    doff = q * ss + k; dstep = dq * ss + dk; dwrap = ss - L;
  }
  MGQP_MEMBER void next() {
    q += dq; k += dk; soff += sstep; doff += dstep;
    if (k >= L) { k -= L; q++; soff += swrap; doff += dwrap; }
  }
};

// A tile is NATURAL when it is full and each of its G, g0, CI and ci0 blocks is one contiguous array (the
// fields' natural strides, no compressed CI): the common case, served by loops whose global offsets are
// 32-bit and mostly compile-time.
template <int N>
MGQP_DEV bool ln_tile_is_natural(const LaneView& v, const LaneTileIO& t) {
  return t.cnt == 32 && !t.CIv && t.stride[0] == N * N && t.stride[1] == N && t.stride[4] == (long long)N * v.m &&
         t.stride[5] == v.m;
}

// pull the NEXT tile's G / g0 blocks towards L2 (natural strides assumed: a wrong guess only wastes a prefetch)
MGQP_DEV void ln_prefetch_next(const LaneTileIO& t, int n) {
  if (!t.next_G) return;
  const int lane = wp_lane();
  const char* p = reinterpret_cast<const char*>(t.next_G);
#pragma unroll 1
  for (int o = lane * 128; o < 32 * n * n * 8; o += 32 * 128) MGQP_PREFETCH_NEXT(p + o);
  const char* p0 = reinterpret_cast<const char*>(t.next_g0);
#pragma unroll 1
  for (int o = lane * 128; o < 32 * n * 8; o += 32 * 128) MGQP_PREFETCH_NEXT(p0 + o);
}

// pull the tile's CI / ci0 blocks towards L2 while phases A and S run (phase B reads them)
// ... and its CI / ci0 blocks: phase B of the next tile is this kernel's one pass over the bulk of its input, and
// it finds the lines in L2 instead of waiting on DRAM
MGQP_DEV void ln_prefetch_next_CI(const LaneView& v, const LaneTileIO& t, int n) {
  if (!t.next_CI) return;
  const int lane = wp_lane();
  const char* p = reinterpret_cast<const char*>(t.next_CI);
  const int bytes = 32 * n * v.m * 8;
#pragma unroll 1
  for (int o = lane * 128; o < bytes; o += 32 * 128) MGQP_PREFETCH_L2(p + o);
  const char* p0 = reinterpret_cast<const char*>(t.next_ci0);
#pragma unroll 1
  for (int o = lane * 128; o < 32 * v.m * 8; o += 32 * 128) MGQP_PREFETCH_L2(p0 + o);
}

MGQP_DEV void ln_prefetch_CI(const LaneView& v, const LaneTileIO& t, int n) {
  const int lane = wp_lane();
  const char* p = reinterpret_cast<const char*>(t.in[4]);
  const int bytes = 32 * n * v.m * 8;
#pragma unroll 4
  for (int o = lane * 128; o < bytes; o += 32 * 128) MGQP_PREFETCH_L2(p + o);
  const char* p0 = reinterpret_cast<const char*>(t.in[5]);
#pragma unroll 1
  for (int o = lane * 128; o < 32 * v.m * 8; o += 32 * 128) MGQP_PREFETCH_L2(p0 + o);
}

// ---- phase A: G and g0 of the tile's QPs into the owners' slots (L scratch, g0/y/x scratch).
template <int N>
MGQP_DEV void ln_stage_G(const LaneView& v, double* __restrict__ ws, const LaneTileIO& t, bool natural) {
  const int lane = wp_lane();
  if (natural) {
    {
      constexpr int L = N * N, dq = 32 / L, dk = 32 - dq * L;
      const double* __restrict__ src = t.in[0] + lane;
      int k = lane % L;
      int doff = (lane / L) * v.stride + k;
      const int dstep = dq * v.stride + dk, dwrap = v.stride - L;
      double* dst = ws + v.oLs;
#pragma unroll 7
      for (int r = 0; r < L; r++) {  // 32 * N*N elements = N*N rounds of 32 lanes
        MGQP_CP8(dst + doff, src + 32 * r);   // (taking the shared-window base once, as phase B does: 0.4 % slower here)
        k += dk; doff += dstep;
        if (k >= L) { k -= L; doff += dwrap; }
      }
    }
    {
      constexpr int L = N, dq = 32 / L, dk = 32 - dq * L;
      const double* __restrict__ src = t.in[1] + lane;
      int k = lane % L;
      int doff = (lane / L) * v.stride + k;
      const int dstep = dq * v.stride + dk, dwrap = v.stride - L;
      double* dst = ws + v.ov;
#pragma unroll
      for (int r = 0; r < L; r++) {
        MGQP_CP8(dst + doff, src + 32 * r);
        k += dk; doff += dstep;
        if (k >= L) { k -= L; doff += dwrap; }
      }
    }
  } else {
    {
      const double* __restrict__ G = t.in[0];
      LaneWalk w;
      w.init(lane, N * N, t.stride[0], v.stride);
      double* dst = ws + v.oLs;
#pragma unroll 1
      for (int r = 0; r < N * N; r++) {
        MGQP_CP8(dst + w.doff, G + (w.q < t.cnt ? w.soff : (long long)w.k));
        w.next();
      }
    }
    {
      const double* __restrict__ g0 = t.in[1];
      LaneWalk w;
      w.init(lane, N, t.stride[1], v.stride);
      double* dst = ws + v.ov;
#pragma unroll 1
      for (int r = 0; r < N; r++) {
        MGQP_CP8(dst + w.doff, g0 + (w.q < t.cnt ? w.soff : (long long)w.k));
        w.next();
      }
    }
  }
  MGQP_CP_WAIT();
}

// One inequality column -> its byte.  `mask` has bit j set iff entry j is nonzero; `val` is the ordered sum of
// the column, which IS the single nonzero bit for bit when there is exactly one (v + (+-0.0) == v).
// (-0.0 counts as a zero: every use of a coefficient is a product added to a sum that starts at +0.0, where
// the sign of a zero coefficient cannot show.)
template <int N>
MGQP_DEV unsigned ln_column_byte(unsigned mask, double val) {
  const unsigned row8 = (unsigned)(wp_ffs(mask) - 1) << LN_IDX_SHIFT;  // mask == 0: overwritten below
  unsigned byte = LN_GEN | row8;                                     // one nonzero, not a unit: exact path
  if (val == 1.0) byte = row8;                                       // +x_k
  if (val == -1.0) byte = row8 + (unsigned)(8 * N);                  // -x_k
  if (mask & (mask - 1)) byte = LN_GEN;                              // two or more nonzeros
  if (mask == 0) byte = (unsigned)(2 * N) << LN_IDX_SHIFT;           // empty column: contributes +0.0
  return byte;
}

// ---- phase B: classify the tile's 32*m inequality columns, lane l <-> columns l, l+32, ... of the flat
// (QP, column) enumeration; copy ci0 into the slots.  Every load is one row segment of CI: coalesced.
// Lanes beyond the tile's last QP get empty columns with ci0 = 1 (always satisfied).
template <int N>
MGQP_DEV void ln_stage_CI(const LaneView& v, double* __restrict__ ws, const LaneTileIO& t, bool natural) {
  const int m = v.m;
  if (m <= 0) return;
  const int lane = wp_lane();
  const double* __restrict__ ci0 = t.in[5];
  unsigned char* bytes = reinterpret_cast<unsigned char*>(ws + v.oByte) + v.bcidx;
  unsigned char* gen = reinterpret_cast<unsigned char*>(ws + v.oByte) + v.bgen;
  double* dci0 = ws + v.oci0;
  if (natural) {
    // the tile's CI block is one array: column (q, c), row j sits at q*N*m + j*m + c = e + j*m with
    // e = t + q*(N-1)*m, t = q*m + c the flat column number (lane, lane + 32, ...)
    const double* __restrict__ CI = t.in[4];
    const double* rowp[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
      rowp[j] = CI + j * m;
      MGQP_OPAQUE_PTR(rowp[j]);
    }
    const int dq = 32 / m, dc = 32 - dq * m;
    int q = lane / m, c = lane - q * m;
    unsigned tt = lane, e = lane + q * (N - 1) * m;
    int doff = q * v.stride + c;
    const int dstep = dq * v.stride + dc, dwrap = v.stride - m;
    const unsigned de = 32 + dq * (N - 1) * m, ewrap = (N - 1) * m;
    const unsigned sdci0 = MGQP_SHARED_BASE(dci0);
    (void)sdci0;
    unsigned seen_gen = 0;
    (void)seen_gen;
    constexpr int K = MGQP_LANE_K;               // columns per trip; two trips (2*K*N loads) in flight per lane
#if MGQP_LANE_CI0_LDG
    struct Trip { double b[K][N]; double c0[K]; int c[K], doff[K]; };   // c0: the column's ci0, loaded with its rows
#else
    struct Trip { double b[K][N]; int c[K], doff[K]; unsigned t[K]; };
#endif
    auto load = [&](Trip& T, int r0) {
#pragma unroll
      for (int u = 0; u < K; u++) {
        if (r0 + u < m) {  // warp-uniform
#pragma unroll
          for (int j = 0; j < N; j++) T.b[u][j] = MGQP_LDG(rowp[j] + e);
#if MGQP_LANE_CI0_LDG
          T.c0[u] = MGQP_LDG(ci0 + tt);
          T.c[u] = c; T.doff[u] = doff;
#else
          T.c[u] = c; T.doff[u] = doff; T.t[u] = tt;
#endif
          tt += 32; c += dc; e += de; doff += dstep;
          if (c >= m) { c -= m; e += ewrap; doff += dwrap; }
        }
      }
    };
    auto sort = [&](const Trip& T, int r0) {
#pragma unroll
      for (int u = 0; u < K; u++) {
        if (r0 + u < m) {
          double val = T.b[u][0];
          unsigned mask = 0;
          d_or_if_nonzero(mask, T.b[u][0], 1u);  // NaN counts as nonzero
#pragma unroll
          for (int j = 1; j < N; j++) {
            val = d_add(val, T.b[u][j]);
            d_or_if_nonzero(mask, T.b[u][j], 1u << j);
          }
          const unsigned byte = ln_column_byte<N>(mask, val);
#if MGQP_LANE_CI0_LDG
          dci0[T.doff[u]] = T.c0[u];   // (a cp.async here costs four issue slots: ptxas pads each with three no-ops)
#else
          MGQP_CP8_AT(sdci0, dci0, T.doff[u], ci0 + T.t[u]);
#endif
          bytes[8 * T.doff[u] - 7 * T.c[u]] = (unsigned char)byte;
#if MGQP_LANE_EFF_BYTE
          bytes[8 * T.doff[u] - 7 * T.c[u] + m] = (unsigned char)(byte >> LN_IDX_SHIFT);
#endif
#if MGQP_LANE_GEN_AFTER
          seen_gen |= byte;   // (bit 0 = LN_GEN) the owners' flags are set after the loop, and only if there is one
#else
          if (byte & LN_GEN) gen[8 * (T.doff[u] - T.c[u])] = 1;
#endif
        }
      }
    };
    Trip A, B;
    load(A, 0);
#pragma unroll 1
    for (int r0 = 0; r0 < m; r0 += 2 * K) {
      load(B, r0 + K);
      sort(A, r0);
      load(A, r0 + 2 * K);
      sort(B, r0 + K);
    }
    MGQP_CP_WAIT();
#if MGQP_LANE_GEN_AFTER
    // general columns are the exception (none in anything the reference builds): only a warp that met one pays for
    // telling the owners -- every lane then looks through its own QP's bytes
    if (wp_any((seen_gen & LN_GEN) != 0)) {
      wp_sync();
      const unsigned char* mine = bytes + (size_t)lane * v.stride * 8;
      bool g = false;
#pragma unroll 1
      for (int i = 0; i < m; i++) g = g || (mine[i] & LN_GEN);
      if (g) gen[(size_t)lane * v.stride * 8] = 1;
    }
#endif
    return;
  }
  LaneWalk w;                                  // source offsets: the ci0 block ([q][c], stride sci0)
  w.init(lane, m, t.stride[5], v.stride);
  auto finish = [&](int q, int k, int doff, long long soff, unsigned byte) {
    // slot byte address: q * stride * 8 + k  ==  8 * (doff - k) + k
    const int boff = 8 * (doff - k) + k;
    if (q < t.cnt) {
      MGQP_CP8(dci0 + doff, ci0 + soff);
    } else {
      dci0[doff] = 1.0;
      byte = (unsigned)(2 * N) << LN_IDX_SHIFT;
    }
    bytes[boff] = (unsigned char)byte;
#if MGQP_LANE_EFF_BYTE
    bytes[boff + m] = (unsigned char)(byte >> LN_IDX_SHIFT);
#endif
    if (byte & LN_GEN) gen[8 * (doff - k)] = 1;
  };
  if (t.CIv) {
    // the caller's (or the host packer's) column-compressed form: (value, row) per column
    long long so = w.q * t.sCIv + w.k, ro = w.q * t.sCIr + w.k;
    const long long sstep = w.dq * t.sCIv + w.dk, swrap = t.sCIv - m, rstep = w.dq * t.sCIr + w.dk, rwrap = t.sCIr - m;
#pragma unroll 2
    for (int r = 0; r < m; r++) {
      unsigned byte = 0;
      if (w.q < t.cnt) {
        const double a = MGQP_LDG(t.CIv + so);
        unsigned row = MGQP_LDG(t.CIr + ro);
        row = row < (unsigned)N ? row : (unsigned)(N - 1);  // a corrupt batch stays inside the table
        byte = ln_column_byte<N>(a != 0.0 ? (1u << row) : 0u, a);
      }
      finish(w.q, w.k, w.doff, w.soff, byte);
      const int k0 = w.k;
      w.next();
      so += sstep; ro += rstep;
      if (w.k < k0 + w.dk) { so += swrap; ro += rwrap; }  // the walk wrapped into the next QP
    }
  } else {
    // strided fields or a partial tile: one column per trip
    const double* __restrict__ CI = t.in[4];
    const long long sCI = t.stride[4];
    long long co = w.q * sCI + w.k;              // offset of CI[q][0][k]
    const long long cstep = w.dq * sCI + w.dk, cwrap = sCI - m;
#pragma unroll 1
    for (int r = 0; r < m; r++) {
      unsigned byte = 0;
      if (w.q < t.cnt) {
        const double* src = CI + co;
        double val = MGQP_LDG(src);
        unsigned mask = 0;
        d_or_if_nonzero(mask, val, 1u);
#pragma unroll
        for (int j = 1; j < N; j++) {
          const double bj = MGQP_LDG(src + j * m);
          val = d_add(val, bj);
          d_or_if_nonzero(mask, bj, 1u << j);
        }
        byte = ln_column_byte<N>(mask, val);
      }
      finish(w.q, w.k, w.doff, w.soff, byte);
      const int k0 = w.k;
      w.next();
      co += cstep;
      if (w.k < k0 + w.dk) co += cwrap;
    }
  }
  MGQP_CP_WAIT();
}

// One tile: lane `wp_lane()` owns QP `lane` of the tile (slot S = ws + lane * stride).
template <int N, bool FMA>
MGQP_DEV void gi_lane_tile(const LaneView& v, double* __restrict__ ws, const LaneTileIO& tio, int max_iter,
                           double* x_out, double& f_out, int& status_out, int& iq_out, int& iters_out) {
  const int p = v.p, m = v.m;
  const int lane = wp_lane();
  const int myq = lane < tio.cnt ? lane : 0;
  double* S = ws + lane * v.stride;
  LaneIO io;
  io.G = tio.in[0] + myq * tio.stride[0];
  io.g0 = tio.in[1] + myq * tio.stride[1];
  io.CE = tio.in[2] + myq * tio.stride[2];
  io.ce0 = tio.in[3] + myq * tio.stride[3];
  io.CI = tio.in[4] + myq * tio.stride[4];
  io.ci0 = tio.in[5] + myq * tio.stride[5];
  if (tio.CIv) {
    io.CIv = tio.CIv + myq * tio.sCIv;
    io.CIr = tio.CIr + myq * tio.sCIr;
  }
  double* R = S + v.oR;
  double* sd = S + v.od;  // d; with r and the zero entry behind it: the scan's table [x | -x | 0.0]
  double* sr = S + v.orr;
  double* su = S + v.ou;
  double* sci0 = S + v.oci0;
  signed char* Bt = reinterpret_cast<signed char*>(S + v.oByte);
  signed char* A = Bt + v.bA;
  signed char* Aold = Bt + v.bAold;
  unsigned char* cidx = reinterpret_cast<unsigned char*>(Bt + v.bcidx);
#if MGQP_LANE_EFF_BYTE
  // eff[i]: index into the scan's table [x | -x | 0 | ... | +inf] (doubles from sd) of the term column i contributes --
  // its own entry while the column may be chosen, the +inf entry while it is in the active set or excluded: the
  // scan's inner loop then needs neither the index mask nor the flag test.  Kept in step with cidx's flags.
  unsigned char* eff = reinterpret_cast<unsigned char*>(Bt + v.beff);
  const unsigned eff_inf = (unsigned)(v.oinf - v.od);
  auto eff_refresh = [&](int i) {
    const unsigned c = cidx[i];
    eff[i] = (unsigned char)((c & (LN_EXCL | LN_ACT)) ? eff_inf : (c >> LN_IDX_SHIFT));
  };
#else
  auto eff_refresh = [&](int) {};
#endif
  const double INF = d_inf();

  int status = 0, iq = 0, steps = 0;
  double f = 0.0;
  double x[N];
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = 0.0;

  // ---------------------------------------------------------------- phase A: stage G, g0
  const bool natural = ln_tile_is_natural<N>(v, tio);
#if MGQP_LANE_PREFETCH_NEXT_CI != 2
  if (natural && m > 0) ln_prefetch_CI(v, tio, N);
#endif
  wp_sync();  // the previous tile's results have left the slots
  ln_stage_G<N>(v, ws, tio, natural);
  wp_sync();

  // ---------------------------------------------------------------- set-up
  // Cholesky of G, J = L^-T and x0 = -G^-1 g0 run out of the slot (L in area B, J and g0 -> y -> x in
  // area A), with the outer loops rolled and the independent rows unrolled: a fully unrolled register
  // version was 4 000 instructions of straight-line code and made the kernel instruction-fetch bound.
  double* Ls = S + v.oLs;
  double* vs = S + v.ov;
  double* Js = S + v.oJs;
  double g0r[N];
#pragma unroll
  for (int i = 0; i < N; i++) g0r[i] = vs[i];
  double c1 = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) c1 = d_add(c1, Ls[i * N + i]);

  bool fail = false;
  bool lfin = true;  // every off-diagonal entry of the Cholesky factor is finite
  double bad_pivot = 0.0;
  double c2 = 0.0;
  bool setup_done = false;
#if MGQP_LANE_UNROLLED_SETUP
  {
    // Fast path: the whole set-up unrolled over the compile-time N, in registers, touching only the entries the
    // reference's loops produce something from (56 + 56 + 42 multiply-subtracts at N = 7 instead of the 147 + 147
    // of the row-parallel rolled form below, no loop control, no address arithmetic).  Its 77 divisions have only N
    // distinct divisors -- the pivots -- and run on the split inline path of the division (warp_prims.cuh): one
    // refinement of 1 / l_ii per pivot, three instructions and a range test per quotient, not a branch in the
    // whole block.  Same operations in the same order per element.  The block only decides at its end whether its
    // results stand: a quotient outside the inline path's range (an exact zero in G's factor: sparse G; a tiny or
    // huge entry), a non-positive pivot, or inf / NaN in the factor (a poisoned G) send the tile to the rolled form,
    // whose full loops reproduce the reference in every such case.
    double L[N][N], rl[N];
    bool bad = false, dok = true;
#pragma unroll
    for (int i = 0; i < N; i++) {
      double col[N];  // [0]: the pivot; [j - i]: the numerators of L[j][i]
#pragma unroll
      for (int j = i; j < N; j++) {
        double sum = Ls[i * N + j];
#pragma unroll
        for (int k = i - 1; k >= 0; k--) sum = msub<FMA>(sum, L[i][k], L[j][k]);
        col[j - i] = sum;
      }
      bad = bad || (col[0] <= 0.0);
      L[i][i] = d_sqrt(col[0]);
      rl[i] = d_rcp_refined(L[i][i]);
      lfin = lfin && (L[i][i] < INF);
#pragma unroll
      for (int j = i + 1; j < N; j++) {
        L[j][i] = d_div_rcp(col[j - i], L[i][i], rl[i], dok);
        lfin = lfin && (d_abs(L[j][i]) < INF);
      }
    }
    // J = L^-T, row i = forward_elimination(L, e_i); entries left of the diagonal are exactly +0.0 (finite factor)
    double Jd[N];  // the diagonal, for c2
#pragma unroll
    for (int i = 0; i < N; i++) {
      double Ji[N];
#pragma unroll
      for (int r = 0; r < N; r++) {
        if (r < i) {
          Ji[r] = 0.0;
        } else {
          double acc = (r == i) ? 1.0 : 0.0;
#pragma unroll
          for (int j = i; j < r; j++) acc = msub<FMA>(acc, L[r][j], Ji[j]);
          Ji[r] = d_div_rcp(acc, L[r][r], rl[r], dok);
        }
        Js[i * N + r] = Ji[r];
      }
      Jd[i] = Ji[i];
    }
    // x = -G^-1 g0: forward then backward elimination
    double y[N], xb[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
      double a = g0r[i];
#pragma unroll
      for (int j = 0; j < i; j++) a = msub<FMA>(a, L[i][j], y[j]);
      y[i] = d_div_rcp(a, L[i][i], rl[i], dok);
    }
#pragma unroll
    for (int i = N - 1; i >= 0; i--) {
      double a = y[i];
#pragma unroll
      for (int j = i + 1; j < N; j++) a = msub<FMA>(a, L[j][i], xb[j]);
      xb[i] = d_div_rcp(a, L[i][i], rl[i], dok);
    }
    if (dok && lfin && !bad) {
      setup_done = true;
#pragma unroll
      for (int i = 0; i < N; i++) c2 = d_add(c2, Jd[i]);
      double fs = 0.0;
#pragma unroll
      for (int i = 0; i < N; i++) {
        x[i] = -xb[i];
        fs = mad<FMA>(g0r[i], x[i], fs);
      }
      f = d_mul(0.5, fs);
    } else {
      lfin = true;  // the rolled form below starts over
    }
  }
#endif
  if (!setup_done) {
#pragma unroll 1
  for (int i = 0; i < N && !fail; i++) {
    double sum[N];
#pragma unroll
    for (int j = 0; j < N; j++) sum[j] = Ls[i * N + j];
#pragma unroll 1
    for (int k = i - 1; k >= 0; k--) {
      const double lik = Ls[i * N + k];
      // (all N columns, not only j >= i: the sums of the columns left of the diagonal are never used,
      // and a select per column costs more than the wasted multiply-subtract)
#pragma unroll
      for (int j = 0; j < N; j++) sum[j] = msub<FMA>(sum[j], lik, Ls[j * N + k]);
    }
    double piv = 0.0;
#pragma unroll
    for (int j = 0; j < N; j++)
      if (j == i) piv = sum[j];
    if (piv <= 0.0) {
      fail = true;
      bad_pivot = piv;
    } else {
      const double dii = d_sqrt(piv);
      Ls[i * N + i] = dii;
      lfin = lfin && (dii < INF);
#pragma unroll
      for (int j = 0; j < N; j++)
        if (j > i) {
          const double val = d_div(sum[j], dii);
          Ls[j * N + i] = val;
          Ls[i * N + j] = val;
          lfin = lfin && (d_abs(val) < INF);
        }
    }
  }

  }  // rolled Cholesky
  if (fail) {
    status = 2;
    f = bad_pivot;
  } else if (!setup_done) {
    // J = L^-T: all N rows (row i = forward_elimination(L, e_i)) advance together, column by column
#pragma unroll 1
    for (int r = 0; r < N; r++) {
      double acc[N];
#pragma unroll
      for (int i = 0; i < N; i++) acc[i] = (i == r) ? 1.0 : 0.0;
#pragma unroll 1
      for (int j = 0; j < r; j++) {
        const double lrj = Ls[r * N + j];
#pragma unroll
        for (int i = 0; i < N; i++) acc[i] = msub<FMA>(acc[i], lrj, Js[i * N + j]);
      }
      // Row i of J solves L z = e_i, so z_r = +0.0 for r < i whenever the factor is finite
      // (0.0 - l*0.0 = +0.0, +0.0 / l_rr = +0.0): those 21 divisions are skipped.
      const double lrr = Ls[r * N + r];
#pragma unroll
      for (int i = 0; i < N; i++) {
        if (i <= r || !lfin) Js[i * N + r] = d_div(acc[i], lrr);
        else Js[i * N + r] = 0.0;
      }
    }
#pragma unroll
    for (int i = 0; i < N; i++) c2 = d_add(c2, Js[i * N + i]);
    // x = -G^-1 g0 (cholesky_solve: forward then backward elimination, in place: g0 -> y -> x), f = 0.5 g0.x
    vs[0] = d_div(vs[0], Ls[0]);
#pragma unroll 1
    for (int i = 1; i < N; i++) {
      double a = vs[i];
#pragma unroll 2
      for (int j = 0; j < i; j++) a = msub<FMA>(a, Ls[i * N + j], vs[j]);
      vs[i] = d_div(a, Ls[i * N + i]);
    }
    vs[N - 1] = d_div(vs[N - 1], Ls[(N - 1) * N + (N - 1)]);
#pragma unroll 1
    for (int i = N - 2; i >= 0; i--) {
      double a = vs[i];
#pragma unroll 2
      for (int j = i + 1; j < N; j++) a = msub<FMA>(a, Ls[i * N + j], vs[j]);
      vs[i] = d_div(a, Ls[i * N + i]);
    }
    double fs = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      x[i] = -vs[i];
      fs = mad<FMA>(g0r[i], x[i], fs);
    }
    f = d_mul(0.5, fs);
  }

  // ---------------------------------------------------------------- phase B: inequality columns
  // (after the set-up, whose L scratch it overwrites)
  Bt[v.bgen] = 0;
  wp_sync();
  ln_stage_CI<N>(v, ws, tio, natural);
  if (natural) ln_prefetch_next(tio, N);
#if MGQP_LANE_PREFETCH_NEXT_CI
  if (natural && m > 0) ln_prefetch_next_CI(v, tio, N);
#endif
  wp_sync();
  const bool has_gen = Bt[v.bgen] != 0;  // some inequality column needs the reference's full sum

  if (!fail) {
    // J moves into REGISTERS for the rest of the solve (it waited in its scratch so that phase B had the
    // registers for its loads): every access below has compile-time indices (the Givens sweeps are unrolled
    // over the column index and predicated on the lane's own iq), which keeps J out of shared memory and
    // removes ~100 shared-memory loads from every iteration.
    double Jr[N][N];
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = 0; j < N; j++) Jr[i][j] = Js[i * N + j];
    constexpr int RH = (N - 1) * N - ((N - 1) * (N - 2)) / 2 + N;  // lane_rh_size(N)
#pragma unroll
    for (int i = 0; i < RH; i++) R[i] = 0.0;  // (rolled, this was six instructions per store)
    S[v.ozero] = 0.0;
#if MGQP_LANE_EFF_BYTE
    S[v.oinf] = INF;
#endif

    double R_norm = 1.0;
    const double thr = d_mul(d_mul(d_mul(d_mul((double)m, MGQP_EPS), c1), c2), 100.0);
    bool any_excl = false;  // some column carries LN_EXCL
    int ip = 0, e = 0;                             // e = next equality constraint to add
    double ss = 0.0, s_ip = 0.0;
    bool need_scan = true, l1 = true, done = false;
    double z[N], np[N], dd[N];
#pragma unroll
    for (int i = 0; i < N; i++) { z[i] = 0.0; np[i] = 0.0; dd[i] = 0.0; }
    // The step-1 snapshot (x_old, u_old) is written every outer iteration but read back only on the
    // rare degenerate branch: it lives in thread-local memory on purpose (volatile keeps it there).
    volatile double xold[N], uold[N + 1];
#pragma unroll
    for (int i = 0; i < N; i++) xold[i] = 0.0;
#pragma unroll
    for (int i = 0; i <= N; i++) uold[i] = 0.0;

    // One pass of this loop = one step-direction computation: first the p equality constraints
    // (Appendix A, "for each equality"), then the active-set iteration proper.  add_constraint and
    // delete_constraint each appear exactly once in the code.
#pragma unroll 1
    while (!done) {
      const bool eqm = e < p;
      if (eqm) {
#pragma unroll
        for (int j = 0; j < N; j++) np[j] = MGQP_LDG(io.CE + j * p + e);
      } else if (need_scan) {
        // ------------------------------------------------ step 1 / 2: scan, choose ip
        if (l1) {
          if (any_excl) {  // iaexcl[i] = true for all i
#pragma unroll 1
            for (int i = 0; i < m; i++) {
              cidx[i] &= (unsigned char)~LN_EXCL;
              eff_refresh(i);
            }
            any_excl = false;
          }
          ss = 0.0;
          ip = 0;
        }
        bool finite = true;
#pragma unroll
        for (int j = 0; j < N; j++) {
          sd[j] = x[j];   // the table [x | -x | 0.0] (d and r are dead between add_constraint and the next direction)
          sr[j] = -x[j];
          finite = finite && (d_abs(x[j]) < INF);
        }
        double psi = 0.0;
        const double ss_in = ss;
        if (!has_gen && finite) {
          // every column is +-e_k or empty: s_i = +-x[k] + ci0_i equals the reference's N-term sum wherever a
          // comparison can see it (the two can differ in the sign of a zero only)
          const char* tab = reinterpret_cast<const char*>(sd);
          // three interleaved searches (columns i = 0, 1, 2 mod 3), merged below: the compare-select chain of a
          // single running minimum is one long dependency, three of them overlap
          double s0 = ss, s1 = ss, s2 = ss;
          int i0 = ip, i1 = ip, i2 = ip;
          int i = 0;
#pragma unroll 2
#if MGQP_LANE_EFF_BYTE
          // (a column that may not be chosen reads +inf: +inf + ci0 is +inf or NaN, never below the minimum)
          (void)tab;
          for (; i + 2 < m; i += 3) {
            const double sa = d_add(sd[eff[i]], sci0[i]);
            const double sb = d_add(sd[eff[i + 1]], sci0[i + 1]);
            const double sc = d_add(sd[eff[i + 2]], sci0[i + 2]);
            if (sa < s0) { s0 = sa; i0 = i; }
            if (sb < s1) { s1 = sb; i1 = i + 1; }
            if (sc < s2) { s2 = sc; i2 = i + 2; }
          }
          for (; i < m; i++) {
            const double sa = d_add(sd[eff[i]], sci0[i]);
            if (sa < s0) { s0 = sa; i0 = i; }
          }
#else
          for (; i + 2 < m; i += 3) {
            const unsigned ca = cidx[i], cb = cidx[i + 1], cc = cidx[i + 2];
            const double sa = d_add(*reinterpret_cast<const double*>(tab + (ca & LN_IDX)), sci0[i]);
            const double sb = d_add(*reinterpret_cast<const double*>(tab + (cb & LN_IDX)), sci0[i + 1]);
            const double sc = d_add(*reinterpret_cast<const double*>(tab + (cc & LN_IDX)), sci0[i + 2]);
            if (sa < s0 && !(ca & (LN_EXCL | LN_ACT))) { s0 = sa; i0 = i; }
            if (sb < s1 && !(cb & (LN_EXCL | LN_ACT))) { s1 = sb; i1 = i + 1; }
            if (sc < s2 && !(cc & (LN_EXCL | LN_ACT))) { s2 = sc; i2 = i + 2; }
          }
          for (; i < m; i++) {
            const unsigned ca = cidx[i];
            const double sa = d_add(*reinterpret_cast<const double*>(tab + (ca & LN_IDX)), sci0[i]);
            if (sa < s0 && !(ca & (LN_EXCL | LN_ACT))) { s0 = sa; i0 = i; }
          }
#endif
          // the reference keeps the FIRST index of the strict minimum: smaller value wins, equal values -> smaller
          // index (a search that found nothing still holds the incoming (ss, ip) and loses every comparison)
          if (s1 < s0 || (s1 == s0 && s1 < ss && i1 < i0)) { s0 = s1; i0 = i1; }
          if (s2 < s0 || (s2 == s0 && s2 < ss && i2 < i0)) { s0 = s2; i0 = i2; }
          ss = s0;
          ip = i0;
          // psi = sum_i min(0, s_i) is only ever compared with thr (step 1: |psi| <= thr -> return f).  It is a
          // sum of negative terms, so psi <= each of them (rounding is monotonic), in particular psi <= ss:
          // whenever ss < -thr the test fails without psi being known.  And with ss >= 0 the outcome does not
          // depend on it either: the reference returns the same f from step 1 or, a few lines later, from
          // step 2.  Only a scan whose most violated eligible constraint lies within thr of feasibility
          // needs the sum itself, in the reference's order.
          if (l1 && ss < 0.0 && !(ss < -thr)) {
#pragma unroll 1
            for (int i = 0; i < m; i++) {
              const unsigned c = cidx[i];
              const double sum = d_add(*reinterpret_cast<const double*>(tab + (c & LN_IDX)), sci0[i]);
              d_acc_neg(psi, sum);  // psi += min(0, sum); a NaN sum adds nothing, as in the reference
            }
          } else {
            psi = -INF;  // stands for "fails the test" (with ss >= 0 the scan ends at step 2 instead)
          }
        } else {
#pragma unroll 1
          for (int i = 0; i < m; i++) {
            const unsigned c = cidx[i];
            const int k = (int)((c & LN_IDX) >> LN_IDX_SHIFT);
            double sum;
            if (!(c & LN_GEN) && finite) {
              sum = sd[k];
            } else if (!(c & LN_GEN) || io.CIv) {
              // non-finite x, or a non-unit coefficient that came in compressed form: the reference's full sum
              // with the column rebuilt from its (row, value) form -- a zero coefficient times a non-finite
              // x[j] is a NaN term there, too
              double a;
              int row;
              if (c & LN_GEN) { a = MGQP_LDG(io.CIv + i); row = k; }
              else { a = (k < N) ? 1.0 : ((k < 2 * N) ? -1.0 : 0.0); row = (k < N) ? k : ((k < 2 * N) ? k - N : 0); }
              sum = 0.0;
#pragma unroll
              for (int j = 0; j < N; j++) sum = mad<FMA>((j == row) ? a : 0.0, x[j], sum);
            } else {
              sum = 0.0;
#pragma unroll
              for (int j = 0; j < N; j++) sum = mad<FMA>(MGQP_LDG(io.CI + j * m + i), x[j], sum);
            }
            sum = d_add(sum, sci0[i]);
            d_acc_neg(psi, sum);
            if (sum < ss && !(c & (LN_EXCL | LN_ACT))) {
              ss = sum;
              ip = i;
            }
          }
        }
        const bool picked = ss < ss_in;  // ss only ever moves down, strictly
        if (picked) s_ip = ss;
        if (l1) {
          if (d_abs(psi) <= thr || !(ss < 0.0)) {  // (ss >= 0: step 2 would return the same f; skip the snapshot)
            done = true;
          } else {
#pragma unroll
            for (int i = 0; i <= N; i++)
              if (i < iq) {
                uold[i] = su[i];
                Aold[i] = A[i];
              }
#pragma unroll
            for (int i = 0; i < N; i++) xold[i] = x[i];
          }
        }
        if (!done) {
          if (ss >= 0.0) {
            done = true;
          } else {
            const unsigned c = cidx[ip];
            const int k = (int)((c & LN_IDX) >> LN_IDX_SHIFT);
            if ((c & LN_GEN) && !io.CIv) {
#pragma unroll
              for (int j = 0; j < N; j++) np[j] = MGQP_LDG(io.CI + j * m + ip);
            } else {
              double a;
              int row;
              if (c & LN_GEN) { a = MGQP_LDG(io.CIv + ip); row = k; }
              else { a = (k < N) ? 1.0 : ((k < 2 * N) ? -1.0 : 0.0); row = (k < N) ? k : ((k < 2 * N) ? k - N : 0); }
#pragma unroll
              for (int j = 0; j < N; j++) np[j] = (j == row) ? a : 0.0;  // a == +0 for an empty column
            }
            su[iq] = 0.0;
            A[iq] = (signed char)ip;
            need_scan = false;
          }
        }
        if (done) break;
      }
      if (!eqm) {
        steps++;
        if (steps > max_iter) {
          status = 4;
          break;
        }
      }

      // ------------------------------------------------ step 2a: d = J^T np, z = J[:, iq:] d[iq:], r = R^-1 d
#pragma unroll
      for (int i = 0; i < N; i++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < N; j++) s = mad<FMA>(Jr[j][i], np[j], s);
        dd[i] = s;
        sd[i] = s;
      }
#pragma unroll
      for (int i = 0; i < N; i++) z[i] = 0.0;
#pragma unroll
      for (int j = 0; j < N; j++) {
        if (j >= iq) {  // a branch, not a select: columns below the warp's smallest iq cost nothing
#pragma unroll
          for (int i = 0; i < N; i++) z[i] = mad<FMA>(Jr[i][j], dd[j], z[i]);
        }
      }
      {
        // row i of the packed R starts at MGQP_RH(i, 0) = i*N - i(i-1)/2: a running offset instead of the index
        // arithmetic per element
        int row = MGQP_RH(iq, 0);
#pragma unroll 1
        for (int i = iq - 1; i >= 0; i--) {
          row -= N - i;  // MGQP_RH(i + 1, 0) - MGQP_RH(i, 0) = N - i
          const double* Ri = R + row;
          double s = 0.0;
#pragma unroll 1
          for (int j = i + 1; j < iq; j++) s = mad<FMA>(Ri[j], sr[j], s);
          sr[i] = d_div(d_sub(sd[i], s), Ri[i]);
        }
      }
      double zz = 0.0, znp = 0.0;
#pragma unroll
      for (int i = 0; i < N; i++) {
        zz = mad<FMA>(z[i], z[i], zz);
        znp = mad<FMA>(z[i], np[i], znp);
      }

      // ------------------------------------------------ step 2b / 2c: step lengths, take the step
      bool do_add = false, do_del = false, partial = false;
      int del_l = 0;
      if (eqm) {
        double npx = 0.0;
#pragma unroll
        for (int i = 0; i < N; i++) npx = mad<FMA>(np[i], x[i], npx);
        double t2 = 0.0;
        if (d_abs(zz) > MGQP_EPS) t2 = d_div(d_sub(-npx, MGQP_LDG(io.ce0 + e)), znp);
#pragma unroll
        for (int k = 0; k < N; k++) x[k] = d_add(x[k], d_mul(t2, z[k]));
        su[iq] = t2;
#pragma unroll 1
        for (int k = 0; k < iq; k++) su[k] = d_sub(su[k], d_mul(t2, sr[k]));
        f = d_add(f, d_mul(d_mul(0.5, d_mul(t2, t2)), znp));
        A[iq] = (signed char)(-e - 1);  // A[e]: iq == e while the equalities are being added
        do_add = true;
      } else {
        int l = 0;
        double t1 = INF;
#pragma unroll 1
        for (int k = p; k < iq; k++) {
          if (sr[k] > 0.0) {
            const double q = d_div(su[k], sr[k]);
            if (q < t1) {
              t1 = q;
              l = A[k];
            }
          }
        }
        double t2;
        if (d_abs(zz) > MGQP_EPS) {
          t2 = d_div(-s_ip, znp);
          if (t2 < 0.0) t2 = INF;
        } else {
          t2 = INF;
        }
        const double t = (t2 < t1) ? t2 : t1;
        if (t >= INF) {
          status = 1;
          f = INF;
          break;
        }
        if (t2 >= INF) {
          // dual step only: drop constraint l, stay on ip
#pragma unroll 1
          for (int k = 0; k < iq; k++) su[k] = d_sub(su[k], d_mul(t, sr[k]));
          su[iq] = d_add(su[iq], t);
          cidx[l] &= (unsigned char)~LN_ACT;  // iai[l] = l
          eff_refresh(l);
          do_del = true;
          del_l = l;
        } else {
#pragma unroll
          for (int k = 0; k < N; k++) x[k] = d_add(x[k], d_mul(t, z[k]));
          f = d_add(f, d_mul(d_mul(t, znp), d_add(d_mul(0.5, t), su[iq])));
#pragma unroll 1
          for (int k = 0; k < iq; k++) su[k] = d_sub(su[k], d_mul(t, sr[k]));
          su[iq] = d_add(su[iq], t);
          if (d_abs(d_sub(t, t2)) < MGQP_EPS) {
            do_add = true;  // full step
          } else {
            cidx[l] &= (unsigned char)~LN_ACT;  // partial step: drop l (iai[l] = l), stay on ip
            eff_refresh(l);
            do_del = true;
            del_l = l;
            partial = true;
          }
        }
      }

      // ------------------------------------------------ add_constraint (Appendix A)
      bool degenerate = false;
      if (do_add) {
#pragma unroll
#if MGQP_LANE_SPLIT_SWEEP
        // The sweep's rotations j = N-1 .. iq+1 chain through ONE number only: rotation j-1 starts from
        // distance(d[j-2], |diag_j|).  Pass 1 runs that chain (one division and one square root per rotation);
        // the three remaining divisions of every rotation (c, s, s / (1 + c)) do not feed it, and pass 2 computes
        // them for a few rotations side by side on the split inline path of the division (warp_prims.cuh) --
        // out of a call per rotation with four divisions in a row they were a tenth of the kernel's time
        // (profiles/r02_notes.md).  Same operations on the same operands: a quotient's sign is separated from its
        // magnitude where the numerator's sign is only known later (x / h == -((-x) / h) exactly).
        // [j]: the rotation's hypotenuse, its second operand (sign pending if it is a diag) -- parked in the slot's d
        // and r vectors (dead until the sweep's result is stored there): the registers are all taken by J
        double* gh = sr;
        double* gb = sd;
        bool gact[N], gdg[N];   // [j]: rotation happens; its second operand is the previous rotation's diag
        {
          double cur = dd[N - 1];
          bool cur_dg = false;
#pragma unroll
          for (int j = N - 1; j >= 1; j--) {
            gact[j] = false;
            gh[j] = 1.0;
            gb[j] = cur;
            gdg[j] = cur_dg;
            if (j >= iq + 1) {
              const double h = ln_distance_ni(dd[j - 1], cur);
              gh[j] = h;
              if (!(d_abs(h) < MGQP_EPS)) {
                gact[j] = true;
                cur = h;
                cur_dg = true;
              } else {  // the reference `continue`s: d[j-1] stays what it was
                cur = dd[j - 1];
                cur_dg = false;
              }
            }
          }
        }
        constexpr int GR = MGQP_LANE_SWEEP_GROUP;  // rotations whose divisions run side by side
        bool prev_neg = false;  // the previous rotation's diag is -h
#pragma unroll
        for (int j0 = N - 1; j0 >= 1; j0 -= GR) {
          if (j0 >= iq + 1) {
            double gc[GR], gs[GR], gx[GR], gd[GR];
            bool ok = true;
#pragma unroll
            for (int u = 0; u < GR; u++) {
              const int j = j0 - u;
              if (j >= 1) {
                const bool on = gact[j];
                const double hn = on ? gh[j] : 1.0;
                const double r = d_rcp_refined(hn);
                gc[u] = d_div_rcp(on ? dd[j - 1] : 1.0, hn, r, ok);
                gs[u] = d_div_rcp(on ? d_abs(gb[j]) : 1.0, hn, r, ok);
              }
            }
#pragma unroll
            for (int u = 0; u < GR; u++) {
              const int j = j0 - u;
              if (j >= 1) {
                const bool bneg = gdg[j] ? prev_neg : d_signbit(gb[j]);
                if (bneg) gs[u] = -gs[u];
                const bool flip = gc[u] < 0.0;
                if (flip) {
                  gc[u] = -gc[u];
                  gs[u] = -gs[u];
                }
                gd[u] = flip ? -gh[j] : gh[j];
                if (gact[j]) prev_neg = flip;
              }
            }
#pragma unroll
            for (int u = 0; u < GR; u++) {
              const int j = j0 - u;
              if (j >= 1) {
                const bool on = gact[j];
                const double den = on ? d_add(1.0, gc[u]) : 1.0;
                const double q = d_div_rcp(on ? d_abs(gs[u]) : 1.0, den, d_rcp_refined(den), ok);
                gx[u] = d_signbit(gs[u]) ? -q : q;
              }
            }
            if (ok) {
#pragma unroll
              for (int u = 0; u < GR; u++) {
                const int j = j0 - u;
                if (j >= 1) {
                  if (gact[j]) {
                    dd[j] = 0.0;
                    dd[j - 1] = gd[u];
#pragma unroll
                    for (int k = 0; k < N; k++) {
                      const double t1 = Jr[k][j - 1], t2 = Jr[k][j];
                      const double n1 = d_add(d_mul(t1, gc[u]), d_mul(t2, gs[u]));
                      Jr[k][j - 1] = n1;
                      Jr[k][j] = d_sub(d_mul(gx[u], d_add(t1, n1)), t2);
                    }
                  }
                }
              }
            } else {
              // some quotient of the group needs the division's general routine (an exact zero in d, as sparse
              // problems have them; a tiny or huge operand): the rotations one by one, as the reference orders them
#pragma unroll
              for (int u = 0; u < GR; u++) {
                const int j = j0 - u;
                if (j >= 1) {
                  if (j >= iq + 1) {
                    const LaneGivens g = ln_givens(dd[j - 1], dd[j]);
                    if (g.valid) {
                      dd[j] = 0.0;
                      dd[j - 1] = g.diag;
#pragma unroll
                      for (int k = 0; k < N; k++) {
                        const double t1 = Jr[k][j - 1], t2 = Jr[k][j];
                        const double n1 = d_add(d_mul(t1, g.c), d_mul(t2, g.s));
                        Jr[k][j - 1] = n1;
                        Jr[k][j] = d_sub(d_mul(g.xny, d_add(t1, n1)), t2);
                      }
                      prev_neg = g.diag < 0.0;
                    }
                  }
                }
              }
            }
          }
        }
#elif MGQP_LANE_GIVENS_CHAIN
        double hcur = 1.0;
        if (N >= 2 && N - 1 >= iq + 1) hcur = ln_distance_ni(dd[N >= 2 ? N - 2 : 0], dd[N - 1]);
        for (int j = N - 1; j >= 1; j--) {
          if (j >= iq + 1) {
            // (a lane whose sweep ends here hands the call a harmless next operand)
            const LaneGivensNext g = ln_givens_next(dd[j - 1], dd[j], hcur, (j - 1 >= iq + 1) ? dd[j >= 2 ? j - 2 : 0] : 1.0);
            hcur = g.hnext;
            if (g.valid) {
              dd[j] = 0.0;
              dd[j - 1] = g.diag;
#pragma unroll
              for (int k = 0; k < N; k++) {
                const double t1 = Jr[k][j - 1], t2 = Jr[k][j];
                const double n1 = d_add(d_mul(t1, g.c), d_mul(t2, g.s));
                Jr[k][j - 1] = n1;
                Jr[k][j] = d_sub(d_mul(g.xny, d_add(t1, n1)), t2);
              }
            }
          }
        }
#else
        for (int j = N - 1; j >= 1; j--) {
          if (j >= iq + 1) {
            const LaneGivens g = ln_givens(dd[j - 1], dd[j]);
            if (g.valid) {
              dd[j] = 0.0;
              dd[j - 1] = g.diag;
#pragma unroll
              for (int k = 0; k < N; k++) {
                const double t1 = Jr[k][j - 1], t2 = Jr[k][j];
                const double n1 = d_add(d_mul(t1, g.c), d_mul(t2, g.s));
                Jr[k][j - 1] = n1;
                Jr[k][j] = d_sub(d_mul(g.xny, d_add(t1, n1)), t2);
              }
            }
          }
        }
#endif
#pragma unroll
        for (int i = 0; i < N; i++) sd[i] = dd[i];
        iq++;
#pragma unroll 1
        for (int i = 0, off = iq - 1; i < iq; off += N - i, i++) R[off] = sd[i];  // off = MGQP_RH(i, iq - 1)
        const double dl = d_abs(sd[iq - 1]);
        bool ok = true;
        if (dl <= d_mul(MGQP_EPS, R_norm)) ok = false;
        else R_norm = (R_norm < dl) ? dl : R_norm;
        if (eqm) {
          if (!ok) {
            status = 3;
            break;
          }
          e++;
        } else if (ok) {
          cidx[ip] |= LN_ACT;  // iai[ip] = -1
          eff_refresh(ip);
          need_scan = true;
          l1 = true;
        } else {
          degenerate = true;
          cidx[ip] |= LN_EXCL;  // iaexcl[ip] = false
          eff_refresh(ip);
          any_excl = true;
          do_del = true;
          del_l = ip;
        }
      }

      // ------------------------------------------------ delete_constraint(del_l) (Appendix A)
      if (do_del) {
        int qq = 0;
#pragma unroll 1
        for (int i = p; i < iq; i++)
          if (A[i] == del_l) {
            qq = i;
            break;
          }
#pragma unroll 1
        for (int i = qq; i < iq - 1; i++) {
          A[i] = A[i + 1];
          su[i] = su[i + 1];
          // rows below i+1 are structurally zero in both columns: nothing to move there
#pragma unroll 1
          for (int j = 0; j <= i + 1; j++) R[MGQP_RH(j, i)] = R[MGQP_RH(j, i + 1)];
        }
        A[iq - 1] = A[iq];
        su[iq - 1] = su[iq];
        A[iq] = 0;
        su[iq] = 0.0;
#pragma unroll 1
        for (int j = 0; j < iq; j++) R[MGQP_RH(j, iq - 1)] = 0.0;
        iq--;
        if (iq > 0) {
#pragma unroll
          for (int j = 0; j < N - 1; j++) {
            if (j >= qq && j < iq) {
              const LaneGivens g = ln_givens(R[MGQP_RH(j, j)], R[MGQP_RH(j + 1, j)]);
              if (g.valid) {
                R[MGQP_RH(j + 1, j)] = 0.0;
                R[MGQP_RH(j, j)] = g.diag;
#pragma unroll 1
                for (int k = j + 1; k < iq; k++) {
                  const double t1 = R[MGQP_RH(j, k)], t2 = R[MGQP_RH(j + 1, k)];
                  const double n1 = d_add(d_mul(t1, g.c), d_mul(t2, g.s));
                  R[MGQP_RH(j, k)] = n1;
                  R[MGQP_RH(j + 1, k)] = d_sub(d_mul(g.xny, d_add(t1, n1)), t2);
                }
#pragma unroll
                for (int k = 0; k < N; k++) {
                  const double t1 = Jr[k][j], t2 = Jr[k][j + 1];
                  const double n1 = d_add(d_mul(t1, g.c), d_mul(t2, g.s));
                  Jr[k][j] = n1;
                  Jr[k][j + 1] = d_sub(d_mul(g.xny, d_add(n1, t1)), t2);
                }
              }
            }
          }
        }
      }

      if (degenerate) {
        // restore the state saved at step 1 and go back to step 2 with the stale ss
#pragma unroll 1
        for (int i = 0; i < m; i++) {
          cidx[i] &= (unsigned char)~LN_ACT;  // iai[i] = i
          eff_refresh(i);
        }
#pragma unroll
        for (int i = 0; i <= N; i++)
          if (i >= p && i < iq) {
            A[i] = Aold[i];
            su[i] = uold[i];
            cidx[A[i]] |= LN_ACT;
            eff_refresh(A[i]);
          }
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = xold[i];
        need_scan = true;
        l1 = false;
      } else if (partial) {
        double npx = 0.0;
#pragma unroll
        for (int i = 0; i < N; i++) npx = mad<FMA>(np[i], x[i], npx);
        s_ip = d_add(npx, sci0[ip]);  // s[ip] = CI[:,ip].x + ci0[ip] with the new x
      }
    }
  }

#pragma unroll
  for (int i = 0; i < N; i++) x_out[i] = x[i];
  f_out = f;
  status_out = status;
  iq_out = iq;
  iters_out = steps;
}

}  // namespace mgqp
