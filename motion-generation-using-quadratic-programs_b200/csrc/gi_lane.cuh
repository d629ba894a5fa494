// Goldfarb-Idnani dual active-set solver, ONE QP PER LANE (n <= 8, m <= 64): a warp advances 32
// independent QPs in SIMT fashion.
//
// Replaces the body of the reference's solve_quadprog (lib/QuadProgpp/libquadprog.a: QuadProg++.o,
// prototype include/QuadProgpp/QuadProg++.hh:69-72, behavioural spec SURVEY.md Appendix A) for the
// small joint-space shapes (BASELINE.json configs[1]/[4]: n=7, p=1, m=42).
//
// Why a second kernel family: with n = 7 a lane group of 8 (gi_small.cuh) spends most of its issue
// slots on work that is uniform across the group (divisions, square roots, control) and on ordered
// cross-lane sums; one lane per QP executes each of those once per QP.  The price is shared
// memory: a lane owns a private, bank-skewed slot, so the constraint matrix cannot stay resident
// in dense form.  Instead each inequality column is classified while it streams in from HBM (its
// only read):
//     * at most one nonzero  -> kept as (row, value): joint acceleration / velocity / position /
//       torque bounds, i.e. every inequality the reference itself ever builds
//       (src/mgqp.cpp:1080-1128, limitsMatrix = [-I; +I]);
//     * anything else        -> marked DENSE and re-read from global memory (L2) when scanned.
// The compressed form is used ONLY in the violated-constraint scan, and only while x is finite:
// sum_j CI[j][i]*x[j] with exact-zero coefficients equals 0.0 + v*x[k] bit for bit.  Everything
// else (compute_d, step lengths, the s[ip] refresh) works on the dense column np = CI[:,ip].
//
// Each lane runs the reference's scalar recurrences in the reference's own operation order, so
// results are bit-identical to QuadProg++ by construction (no cross-lane arithmetic at all).
//
// The set-up (Cholesky of G, J = L^-T, x0 = -G^-1 g0) is fully unrolled over the compile-time N and
// lives in registers: G and g0 go from global memory straight into registers; J never leaves them.  The 49 + 21 + 14 IEEE divisions of that phase are independent across rows, which is
// where a lone warp per scheduler gets its instruction-level parallelism from.
//
// Slot layout (doubles; the lane stride is odd so the 16 lanes of a half warp hit 16 distinct
// 8-byte bank pairs whenever they address the same element):
//   R (upper Hessenberg, packed) | d[N] (doubles as the x mirror during scans) | r[N] | u[N+1] |
//   civ[m] | ci0[m] | int8: A[N+1], A_old[N+1], cidx[m]          (J and u_old live in registers)
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
#else
#define MGQP_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
#define MGQP_LDG(p) __ldg(p)
#define MGQP_LDG2(p, a, b) do { const double2 v2_ = __ldg(reinterpret_cast<const double2*>(p)); (a) = v2_.x; (b) = v2_.y; } while (0)
// streaming 16-byte load that is cached in L2 only: each lane touches its own 128-byte lines, and the
// few KB of L1 left beside 224 KB of shared memory cannot hold them until their other sectors are read
#define MGQP_LDG2_STREAM(p, a, b) asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "l"(p))
#endif

#ifndef MGQP_LANE_PQ
#define MGQP_LANE_PQ 1
#endif

namespace mgqp {

struct LaneView {
  int n, p, m;
  int oR, od, orr, ou, ociv, oci0, oByte;               // doubles
  int bA, bAold, bcidx;                                // bytes, relative to (signed char*)(slot + oByte)
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
  v.ou = o;     o += n + 1;
  v.ociv = o;   o += m;
  v.oci0 = o;   o += m;
  // the civ|ci0 area doubles as set-up scratch (L | y | x | g0), the R|d|r|u area as scratch for J
  if (2 * m < n * n + 3 * n) o += n * n + 3 * n - 2 * m;
  v.oByte = o;
  v.bA = 0;
  v.bAold = n + 1;
  v.bcidx = 2 * (n + 1);
  const int bytes = 2 * (n + 1) + m;
  o += (bytes + 7) / 8;
  if ((o & 1) == 0) o++;  // odd stride: the 16 lanes of a half warp land on 16 distinct 8-byte bank pairs
  v.stride = o;
  return v;
}

// per-column byte: ONE (a nonzero has been seen) | DENSE (more than one) | bits 3-5 row of the single
// nonzero, i.e. its byte offset into a vector of doubles | EXCL (iaexcl == false) | ACT (iai == -1).
// A column may be chosen iff byte < LN_EXCL.
enum { LN_ONE = 0x01, LN_DENSE = 0x02, LN_ROW = 0x38, LN_ROW_SHIFT = 3, LN_EXCL = 0x40, LN_ACT = 0x80 };

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

// (c, s) -> normalised rotation coefficients exactly as add_constraint / delete_constraint derive
// them (Appendix A).  One shared, non-inlined copy: it holds four IEEE divisions and a square root,
// and the unrolled sweeps call it from twelve places.  valid == 0 <=> the reference `continue`s.
struct LaneGivens {
  double c, s, xny, diag;
  int valid;
};
MGQP_DEV_NOINLINE LaneGivens ln_givens(double cc, double ss) {
  LaneGivens g;
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

// Stream this lane's CI (N x m, row-major) and ci0 from global memory -- their only read -- and sort the
// inequality columns into (row, value) form: cidx[i] = flag byte, civ[i] = the single nonzero coefficient
// (+-0.0 for an empty column), sci0[i] = ci0[i].  Returns true iff some column has more than one nonzero.
// PQ = 1: with two or three pairs per chunk (or a three-deep ring) ptxas spills the load buffers, and
// as a separate non-inlined routine the call ABI costs more than it buys (measured, profiles/r01_notes.md).
template <int N>
MGQP_DEV bool ln_classify(const double* __restrict__ CI, const double* __restrict__ ci0, int m,
                                   double* __restrict__ civ, double* __restrict__ sci0, unsigned char* __restrict__ cidx) {
  bool has_dense = false;
  const bool vec = ((m & 1) == 0) && ((reinterpret_cast<unsigned long long>(CI) & 15ull) == 0ull) &&
                   ((reinterpret_cast<unsigned long long>(ci0) & 15ull) == 0ull);
  if (vec) {
    // PQ column pairs (one 16-byte load per row of CI, plus one of ci0, per pair) per chunk, the next
    // chunk's PQ*(N+1) loads in flight while the current one is classified.  Per element: one compare,
    // one add into the column's running sum and one conditional integer add into its (count, row)
    // word -- with at most one nonzero in a column the sum IS that coefficient bit for bit
    // (v + (+-0.0) == v), and count*64 + 8*row is the flag byte.
    constexpr int PQ = MGQP_LANE_PQ;
    const int mh = m >> 1;
    double bufA[PQ][N + 1][2], bufB[PQ][N + 1][2];
    auto load_chunk = [&](double (&b)[PQ][N + 1][2], int c) {
#pragma unroll
      for (int q = 0; q < PQ; q++)
        if (c + q < mh) {
#pragma unroll
          for (int j = 0; j < N; j++) MGQP_LDG2_STREAM(CI + j * m + 2 * (c + q), b[q][j][0], b[q][j][1]);
          MGQP_LDG2_STREAM(ci0 + 2 * (c + q), b[q][N][0], b[q][N][1]);
        }
    };
    auto sort_chunk = [&](const double (&b)[PQ][N + 1][2], int c) {
#pragma unroll
      for (int q = 0; q < PQ; q++)
        if (c + q < mh) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
            double val = b[q][0][h];
            unsigned acc = (b[q][0][h] != 0.0) ? 64u : 0u;  // NaN counts as nonzero
#pragma unroll
            for (int j = 1; j < N; j++) {
              val = d_add(val, b[q][j][h]);
              if (b[q][j][h] != 0.0) acc += 64u + 8u * j;
            }
            const int i = 2 * (c + q) + h;
            unsigned byte = (acc & LN_ROW) | LN_ONE;
            if (acc < 64u) byte = 0;
            if (acc >= 128u) {
              byte = LN_ONE | LN_DENSE;
              has_dense = true;
            }
            cidx[i] = (unsigned char)byte;
            civ[i] = val;
            sci0[i] = b[q][N][h];
          }
        }
    };
    // (All of these loads share one hardware scoreboard -- ptxas' choice -- so the first use of a buffer
    // also waits for the loads issued after it: each trip of this loop exposes one full round trip to L2
    // whatever the order of its four stages.  Ordering fences and a three-deep ring were measured: no gain.)
    load_chunk(bufA, 0);
#pragma unroll 1
    for (int c = 0; c < mh; c += 2 * PQ) {
      load_chunk(bufB, c + PQ);
      sort_chunk(bufA, c);
      load_chunk(bufA, c + 2 * PQ);
      sort_chunk(bufB, c + PQ);
    }
  } else {
    // odd m or unaligned fields: element-wise row sweep
#pragma unroll 1
    for (int i = 0; i < m; i++) {
      cidx[i] = 0;
      civ[i] = 0.0;
      sci0[i] = MGQP_LDG(ci0 + i);
    }
#pragma unroll 1
    for (int j = 0; j < N; j++) {
      const double* row = CI + j * m;
#pragma unroll 4
      for (int i = 0; i < m; i++) {
        const double a = MGQP_LDG(row + i);
        if (a != 0.0) {  // NaN counts as nonzero
          if (!(cidx[i] & LN_ONE)) {
            cidx[i] = (unsigned char)((j << LN_ROW_SHIFT) | LN_ONE);
            civ[i] = a;
          } else {
            cidx[i] |= LN_DENSE;
            has_dense = true;
          }
        }
      }
    }
  }
  return has_dense;
}

struct LaneIO {  // global memory, this lane's QP
  const double* G; const double* g0; const double* CE; const double* ce0; const double* CI; const double* ci0;
  // lane kernel only: column-compressed CI (mgqp_batch::CI_val / CI_row); CI is unused when CIv is set
  const double* CIv = nullptr;
  const unsigned char* CIr = nullptr;
};

// Column-compressed constraints as they come from the caller (or from the host pipeline's packer): the
// (row, value) form the classification would have produced, taken as is.  A row index >= N cannot come
// from a well-formed batch; it is clamped so that a corrupt one stays inside the lane's slot.
template <int N>
MGQP_DEV void ln_take_compressed(const double* __restrict__ CIv, const unsigned char* __restrict__ CIr,
                                 const double* __restrict__ ci0, int m, double* __restrict__ civ,
                                 double* __restrict__ sci0, unsigned char* __restrict__ cidx) {
  const bool vec = ((m & 1) == 0) && ((reinterpret_cast<unsigned long long>(CIv) & 15ull) == 0ull) &&
                   ((reinterpret_cast<unsigned long long>(ci0) & 15ull) == 0ull);
  if (vec) {
#pragma unroll 7
    for (int i = 0; i < m; i += 2) {
      double a0, a1, b0, b1;
      MGQP_LDG2(CIv + i, a0, a1);
      MGQP_LDG2(ci0 + i, b0, b1);
      civ[i] = a0; civ[i + 1] = a1;
      sci0[i] = b0; sci0[i + 1] = b1;
    }
  } else {
#pragma unroll 4
    for (int i = 0; i < m; i++) {
      civ[i] = MGQP_LDG(CIv + i);
      sci0[i] = MGQP_LDG(ci0 + i);
    }
  }
#pragma unroll 4
  for (int i = 0; i < m; i++) {
    unsigned r = MGQP_LDG(CIr + i);
    r = r < (unsigned)N ? r : (unsigned)(N - 1);
    cidx[i] = (unsigned char)((r << LN_ROW_SHIFT) | LN_ONE);
  }
}

template <int N, bool FMA>
MGQP_DEV void gi_lane_solve(const LaneView& v, double* S, const LaneIO& io, int max_iter, double* x_out,
                            double& f_out, int& status_out, int& iq_out, int& iters_out) {
  const int p = v.p, m = v.m;
  double* R = S + v.oR;
  double* sd = S + v.od;  // d; reused as the smem mirror of x while scanning
  double* sr = S + v.orr;
  double* su = S + v.ou;
  double* civ = S + v.ociv;
  double* sci0 = S + v.oci0;
  signed char* Bt = reinterpret_cast<signed char*>(S + v.oByte);
  signed char* A = Bt + v.bA;
  signed char* Aold = Bt + v.bAold;
  unsigned char* cidx = reinterpret_cast<unsigned char*>(Bt + v.bcidx);
  const double INF = d_inf();

  int status = 0, iq = 0, steps = 0;
  double f = 0.0;
  bool has_dense = false;  // some inequality column has more than one nonzero
  double x[N];
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = 0.0;

  // ---------------------------------------------------------------- set-up
  // Cholesky of G, J = L^-T and x0 = -G^-1 g0 run out of shared-memory scratch that the constraint
  // data will occupy later (L | y | x | g0 in the civ/ci0 area, J in the R|d|r|u area), with the
  // outer loops rolled and the independent rows unrolled: a fully unrolled register version was
  // 4 000 instructions of straight-line code and made the kernel instruction-fetch bound.
  // Pull this lane's constraint data towards L2 now; it is consumed right after the set-up (a
  // whole-tile-ahead prefetch was evicted before use and doubled the DRAM traffic, this one is not).
  {
    const char* pc = reinterpret_cast<const char*>(io.CIv ? io.CIv : io.CI);
    const int bytes = (io.CIv ? 1 : N) * m * 8;
#pragma unroll 4
    for (int o = 0; o < bytes; o += 128) MGQP_PREFETCH_L2(pc + o);
    const char* p0 = reinterpret_cast<const char*>(io.ci0);
#pragma unroll 1
    for (int o = 0; o < m * 8; o += 128) MGQP_PREFETCH_L2(p0 + o);
  }
  double* Ls = S + v.ociv;
  double* ys = Ls + N * N;
  double* xs = ys + N;
  double* g0s = xs + N;
  double* Js = S + v.oR;
  // (element-wise on purpose: 16-byte loads at each lane's own parity halve the requests but measured
  // 4 % slower -- these loads hit L1 for the rest of each sector and are issued long before their use)
#pragma unroll
  for (int i = 0; i < N * N; i++) Ls[i] = MGQP_LDG(io.G + i);
#pragma unroll
  for (int i = 0; i < N; i++) g0s[i] = MGQP_LDG(io.g0 + i);
  double c1 = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) c1 = d_add(c1, Ls[i * N + i]);

  bool fail = false;
  bool lfin = true;  // every off-diagonal entry of the Cholesky factor is finite
  double bad_pivot = 0.0;
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

  if (fail) {
    status = 2;
    f = bad_pivot;
  } else {
    // J = L^-T: all N rows (row i = forward_elimination(L, e_i)) advance together, column by column
    double c2 = 0.0;
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
    // x = -G^-1 g0 (cholesky_solve: forward then backward elimination), f = 0.5 g0.x
    ys[0] = d_div(g0s[0], Ls[0]);
#pragma unroll 1
    for (int i = 1; i < N; i++) {
      double a = g0s[i];
#pragma unroll 2
      for (int j = 0; j < i; j++) a = msub<FMA>(a, Ls[i * N + j], ys[j]);
      ys[i] = d_div(a, Ls[i * N + i]);
    }
    xs[N - 1] = d_div(ys[N - 1], Ls[(N - 1) * N + (N - 1)]);
#pragma unroll 1
    for (int i = N - 2; i >= 0; i--) {
      double a = ys[i];
#pragma unroll 2
      for (int j = i + 1; j < N; j++) a = msub<FMA>(a, Ls[i * N + j], xs[j]);
      xs[i] = d_div(a, Ls[i * N + i]);
    }
    double fs = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      x[i] = -xs[i];
      fs = mad<FMA>(g0s[i], x[i], fs);
    }
    f = d_mul(0.5, fs);
    // ---------------------------------------------------------------- inequality columns: classify
    // (after the set-up, whose L | y | x | g0 scratch it overwrites; J waits in its own scratch so that
    // the registers are free for the load pipeline)
    if (m > 0) {
      if (io.CIv) ln_take_compressed<N>(io.CIv, io.CIr, io.ci0, m, civ, sci0, cidx);
      else has_dense = ln_classify<N>(io.CI, io.ci0, m, civ, sci0, cidx);
    }

    // J moves into REGISTERS for the rest of the solve: every access below has compile-time indices
    // (the Givens sweeps are unrolled over the column index and predicated on the lane's own iq),
    // which keeps J out of shared memory (6 resident warps per SM instead of 4) and removes ~100
    // shared-memory loads from every iteration.
    double Jr[N][N];
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = 0; j < N; j++) Jr[i][j] = Js[i * N + j];

#pragma unroll 1
    for (int i = 0; i < lane_rh_size(N); i++) R[i] = 0.0;

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
    // Stores do not stall, and the 30 registers it would take otherwise made the compiler spill
    // loop scalars whose reloads -- L1 is tiny next to 224 KB of shared memory -- cost an L2 round trip.
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
            for (int i = 0; i < m; i++) cidx[i] &= (unsigned char)~LN_EXCL;
            any_excl = false;
          }
          ss = 0.0;
          ip = 0;
        }
        bool finite = true;
#pragma unroll
        for (int j = 0; j < N; j++) {
          sd[j] = x[j];  // x mirror (d is dead between add_constraint and the next direction)
          finite = finite && (d_abs(x[j]) < INF);
        }
        double psi = 0.0;
        const double ss_in = ss;
        if (!has_dense && finite) {
          // every column is (row, value) or empty: s_i = v*x[row] + ci0_i equals the reference's N-term
          // sum wherever a comparison can see it (the two can differ in the sign of a zero only)
#pragma unroll 6
          for (int i = 0; i < m; i++) {
            const unsigned c = cidx[i];
            const double sum = d_add(d_mul(civ[i], sd[(c & LN_ROW) >> LN_ROW_SHIFT]), sci0[i]);
            d_acc_neg(psi, sum);  // psi += min(0, sum); a NaN sum adds nothing, as in the reference
            if (sum < ss && c < LN_EXCL) {
              ss = sum;
              ip = i;
            }
          }
        } else {
#pragma unroll 1
          for (int i = 0; i < m; i++) {
            const unsigned c = cidx[i];
            double sum;
            if (!(c & LN_DENSE) && finite) {
              sum = d_mul(civ[i], sd[(c & LN_ROW) >> LN_ROW_SHIFT]);
            } else if (!(c & LN_DENSE)) {
              // non-finite x: the reference's full sum, with the column rebuilt from its (row, value)
              // form -- a zero coefficient times a non-finite x[j] is a NaN term there, too
              sum = 0.0;
              const int k = (int)((c & LN_ROW) >> LN_ROW_SHIFT);
              const double a = civ[i];
#pragma unroll
              for (int j = 0; j < N; j++) sum = mad<FMA>((j == k) ? a : 0.0, x[j], sum);
            } else {
              sum = 0.0;
#pragma unroll
              for (int j = 0; j < N; j++) sum = mad<FMA>(MGQP_LDG(io.CI + j * m + i), x[j], sum);
            }
            sum = d_add(sum, sci0[i]);
            d_acc_neg(psi, sum);
            if (sum < ss && c < LN_EXCL) {
              ss = sum;
              ip = i;
            }
          }
        }
        const bool picked = ss < ss_in;  // ss only ever moves down, strictly
        if (picked) s_ip = ss;
        if (l1) {
          if (d_abs(psi) <= thr) {
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
            if (c & LN_DENSE) {
#pragma unroll
              for (int j = 0; j < N; j++) np[j] = MGQP_LDG(io.CI + j * m + ip);
            } else {
              const double a = civ[ip];
#pragma unroll
              for (int j = 0; j < N; j++) np[j] = (j == (int)((c & LN_ROW) >> LN_ROW_SHIFT)) ? a : 0.0;  // a == +-0 for an empty column
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
#pragma unroll 1
      for (int i = iq - 1; i >= 0; i--) {
        double s = 0.0;
#pragma unroll 1
        for (int j = i + 1; j < iq; j++) s = mad<FMA>(R[MGQP_RH(i, j)], sr[j], s);
        sr[i] = d_div(d_sub(sd[i], s), R[MGQP_RH(i, i)]);
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
#pragma unroll
        for (int i = 0; i < N; i++) sd[i] = dd[i];
        iq++;
#pragma unroll 1
        for (int i = 0; i < iq; i++) R[MGQP_RH(i, iq - 1)] = sd[i];
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
          need_scan = true;
          l1 = true;
        } else {
          degenerate = true;
          cidx[ip] |= LN_EXCL;  // iaexcl[ip] = false
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
        for (int i = 0; i < m; i++) cidx[i] &= (unsigned char)~LN_ACT;  // iai[i] = i
#pragma unroll
        for (int i = 0; i <= N; i++)
          if (i >= p && i < iq) {
            A[i] = Aold[i];
            su[i] = uold[i];
            cidx[A[i]] |= LN_ACT;
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
