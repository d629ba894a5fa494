// Goldfarb-Idnani dual active-set solver, one QP per GROUP of GW lanes (GW = 8, 16 or 32, n <= GW),
// 32/GW QPs per warp, all state in shared memory and registers.
//
// This is the B200 replacement for the body of the reference's solve_quadprog
// (lib/QuadProgpp/libquadprog.a : QuadProg++.o; prototype include/QuadProgpp/QuadProg++.hh:69-72;
// behavioural spec SURVEY.md Appendix A; helper arithmetic as in
// include/QuadProgpp/eigen/ArrayHelpers.hh:66-99,174-200,239-284).  It is NOT a translation of the
// scalar code: the warp runs its QPs in lock step, one vector element (or one matrix row/column,
// or a strided slice of the inequality constraints) per lane, and every branch of the active-set
// iteration (dual step / full step / partial step / degenerate add) is a predicated section of one
// warp-uniform loop.  What IS kept, deliberately, is the reference's floating-point expression
// order: every dot product is accumulated in ascending index order with one rounding per operation
// (template FMA=false), every tolerance test and tie-break is the reference's, so that the final
// active set, the verdict, x and f come out bit-identical to QuadProg++.  (The single exception is
// the infeasibility sum psi, which is tree-reduced and only recomputed in reference order when it
// lands within a factor 2 of the termination threshold -- the one place its value matters.)
//
// Lane roles inside a group (li = lane index in the group):
//   vectors x, z, d, r, u, A     : element li lives in lane li's registers (d, r, x mirrored in smem)
//   J (n x n, smem, ld = ldJ)    : lane li owns ROW li for the Givens sweeps and update_z, reads
//                                  COLUMN li for compute_d
//   R (n x n, smem, ld = ldR)    : lane li owns row li for the column shift, column li for the sweep
//   inequality i = li + k*GW     : scanned by lane li (k = 0 .. ceil(m/GW)-1); its "active" and
//                                  "excluded" flags (iai / iaexcl in the reference) are bit k of two
//                                  per-lane masks  =>  m <= 32*GW
#pragma once
#include "warp_prims.cuh"

#ifndef MGQP_UNROLL_HOT
#define MGQP_UNROLL_HOT MGQP_UNROLL4  // the four hottest loops (d, z, scan column, Givens row streaming); 2 -> 4: +1 %
#endif

namespace mgqp {

#define MGQP_EPS 2.220446049250313e-16 /* std::numeric_limits<double>::epsilon() */

enum { ST_OK = 0, ST_INFEASIBLE = 1, ST_NOT_PD = 2, ST_EQ_DEPENDENT = 3, ST_ITER_CAP = 4 };

// Shared-memory plan of one warp (offsets in doubles); identical for every warp of a launch.
struct TileView {
  int n, p, m;
  int qpw;                               // QPs per warp = 32 / GW
  int oG, og0, oCE, oce0, oCI, oci0;     // input block, field-major; QP stride = natural size
  int ci_comp, cisz;                     // CI kept column-compressed in shared memory (values | rows); doubles per QP
  int oWork, wstride;                    // work block, QP-major
  int wJ, wR, wx, wd, wr, wred, wgc, wgs, wgx;  // offsets inside one QP's work block
  int ldJ, ldR;
  int total;                             // doubles per warp (multiple of 2)
};

MGQP_HD inline int mgqp_even(int v) { return (v + 1) & ~1; }

// Host-side (and emulator-side) construction of the plan.
// ci_comp: the batch's inequality block arrives column-compressed (mgqp_batch::CI_val / CI_row) and STAYS that way in
// shared memory -- m values and m int rows per QP instead of the dense n x m block, which for the reference's shipped
// shapes (n = 14, m = 28) is a third of a QP's footprint: more warps per SM.
MGQP_HOST inline TileView make_tile_view(int n, int p, int m, int gw, int ci_comp = 0) {
  TileView tv;
  tv.n = n; tv.p = p; tv.m = m;
  tv.qpw = 32 / gw;
  tv.ci_comp = ci_comp;
  tv.cisz = ci_comp ? m + (m + 1) / 2 : n * m;
  int o = 0;
  // every field offset is kept even (16-byte aligned) so that a whole tile of a field can be
  // fetched with one bulk copy
  tv.oG = o;   o = mgqp_even(o + tv.qpw * n * n);
  tv.og0 = o;  o = mgqp_even(o + tv.qpw * n);
  tv.oCE = o;  o = mgqp_even(o + tv.qpw * n * p);
  tv.oce0 = o; o = mgqp_even(o + tv.qpw * p);
  tv.oCI = o;  o = mgqp_even(o + tv.qpw * tv.cisz);
  tv.oci0 = o; o = mgqp_even(o + tv.qpw * m);
  tv.oWork = o;
  tv.ldJ = n | 1;  // odd leading dimension: row-strided 64-bit accesses stay bank-conflict free
  tv.ldR = n | 1;
  int w = 0;
  tv.wJ = w;   w += n * tv.ldJ;
  tv.wR = w;   w += n * tv.ldR;
  tv.wx = w;   w += n;
  tv.wd = w;   w += n;
  tv.wr = w;   w += n;
  tv.wred = w; w += gw;
  tv.wgc = w;  w += gw;  // Givens coefficients of the add_constraint sweep, one rotation per lane
  tv.wgs = w;  w += gw;
  tv.wgx = w;  w += gw;
  tv.wstride = mgqp_even(w) + 2;  // +2 doubles skews consecutive QPs across banks
  tv.total = mgqp_even(o + tv.qpw * tv.wstride);
  return tv;
}

struct TileOut {   // global-memory outputs, already offset to the first QP of the tile
  double* x; double* f; int* status; int* iq; int* A; int* iters;
  // solveNextStep's retry (src/mgqp.cpp:717-735) answered from THIS solve where that is exact: the second call
  // runs without inequalities on the G the first call left behind, its Cholesky factor.  With G == I (the
  // reference's own caller, :672) the factor is G itself, so the retry repeats this solve's equality phase
  // operation for operation and returns the state right after it.  When set, a QP of an all-identity warp whose
  // answer is NaN / +inf is reported as that state with MGQP_STATUS_RETRIED (= 16) instead.
  bool fuse_retry = false;
};

#ifndef MGQP_SMALL_SPLIT_DIV
#define MGQP_SMALL_SPLIT_DIV 1
#endif

// sqrt(a^2+b^2) exactly as QuadProg++'s `distance` (weak symbol in QuadProg++.o)
MGQP_DEV double gi_distance(double a, double b) {
  // One division and one square root whichever argument is the larger: the two branches of the reference
  // are the same formula with the arguments swapped, and lanes that disagree on the comparison would
  // otherwise run both.  Equal magnitudes -- and the reference's fall-through for unordered ones -- return
  // a1*sqrt(2) up front: sparse problems (G = I, bound constraints) meet distance(0, 0) all the time.
  const double a1 = d_abs(a), b1 = d_abs(b);
  const bool agt = a1 > b1;
  if (!agt && !(b1 > a1)) return d_mul(a1, d_sqrt(2.0));
  const double hi = agt ? a1 : b1, lo = agt ? b1 : a1;
#if MGQP_SMALL_SPLIT_DIV
  // the chain of hypotenuses is serial and every lane of the group repeats it: no convergence region on the way
  // (split inline paths of the division and the square root); zeros and out-of-range operands take the plain form
  {
    bool ok = true;
    const double t = d_div_rcp(lo, hi, d_rcp_refined(hi), ok);
    const double h = d_mul(hi, d_sqrt_split(d_add(1.0, d_mul(t, t)), ok));
    if (ok) return h;
  }
#endif
  const double t = d_div_z(lo, hi);  // distance(0, h) is the common case on sparse problems
  return d_mul(hi, d_sqrt(d_add(1.0, d_mul(t, t))));
}

// One Givens rotation's coefficients from (cc, ss) as add_constraint / delete_constraint derive
// them.  Returns false when the reference `continue`s (|h| < eps).  diag = the value the pivot
// entry takes (+-h).
MGQP_DEV_NOINLINE bool gi_givens(double& cc, double& ss, double& xny, double& diag) {
  const double h = gi_distance(cc, ss);
  if (d_abs(h) < MGQP_EPS) return false;
  ss = d_div(ss, h);
  cc = d_div(cc, h);
  if (cc < 0.0) {
    cc = -cc;
    ss = -ss;
    diag = -h;
  } else {
    diag = h;
  }
  xny = d_div(ss, d_add(1.0, cc));
  return true;
}

template <int GW, bool FMA>
MGQP_DEV void gi_solve_tile(const TileView& tv, double* ws, int cnt, int max_iter, const TileOut& out) {
  constexpr int LOGG = (GW == 8) ? 3 : (GW == 16) ? 4 : 5;
  constexpr int NONE = 0x7fffffff;
  const int lane = wp_lane();
  const int g = lane >> LOGG, li = lane & (GW - 1);
  const int n = tv.n, p = tv.p, m = tv.m;
  const int ldJ = tv.ldJ, ldR = tv.ldR;
  const bool in_n = li < n;
  const double INF = d_inf();

  double* sG = ws + tv.oG + g * n * n;
  const double* sg0 = ws + tv.og0 + g * n;
  const double* sCE = ws + tv.oCE + g * n * p;
  const double* sce0 = ws + tv.oce0 + g * p;
  const double* sCI = ws + tv.oCI + g * tv.cisz;  // dense: [n][m]; compressed: value[m] | row[m] (int)
  const bool comp = tv.ci_comp != 0;
  const int* sCIr = reinterpret_cast<const int*>(sCI + m);
  const double* sci0 = ws + tv.oci0 + g * m;
  double* wk = ws + tv.oWork + g * tv.wstride;
  double* sJ = wk + tv.wJ;
  double* sR = wk + tv.wR;
  double* sx = wk + tv.wx;
  double* sd = wk + tv.wd;
  double* sr = wk + tv.wr;
  double* sred = wk + tv.wred;
  double* sgc = wk + tv.wgc;
  double* sgs = wk + tv.wgs;
  double* sgx = wk + tv.wgx;
  // Column i of a compressed block times x, as the reference's n-term sum from +0.0 comes out: the zero terms leave
  // a finite running sum alone (and turn -0.0 into +0.0), but 0 * inf = NaN -- so a non-finite x entry anywhere
  // but in the column's own row poisons the sum.  nbad / kbad: number of non-finite entries of x, index of the first.
  auto comp_dot = [&](int i, int nbad, int kbad) -> double {
    const int k = sCIr[i];
    const double prod = d_add(0.0, d_mul(sCI[i], sx[k]));
    return (nbad == 0 || (nbad == 1 && k == kbad)) ? prod : d_sub(INF, INF);
  };

  // ordered (ascending-index) sum of one value per lane, replicated to the whole group
  auto osum = [&](double v) -> double {
    if (in_n) sred[li] = v;
    wp_sync();
    double acc = 0.0;
    MGQP_UNROLL1
    for (int i = 0; i < n; i++) acc = d_add(acc, sred[i]);
    wp_sync();
    return acc;
  };

  // two / three such sums in one pass (the Givens-coefficient rows double as scratch: they are only
  // live inside add_constraint): one pair of barriers and one loop instead of three, each sum still
  // accumulated in ascending index order
  auto osum2 = [&](double v0, double v1, double& r0, double& r1) {
    if (in_n) {
      sred[li] = v0;
      sgc[li] = v1;
    }
    wp_sync();
    double a0 = 0.0, a1 = 0.0;
    MGQP_UNROLL2
    for (int i = 0; i < n; i++) {
      a0 = d_add(a0, sred[i]);
      a1 = d_add(a1, sgc[i]);
    }
    wp_sync();
    r0 = a0;
    r1 = a1;
  };
  auto osum3 = [&](double v0, double v1, double v2, double& r0, double& r1, double& r2) {
    if (in_n) {
      sred[li] = v0;
      sgc[li] = v1;
      sgs[li] = v2;
    }
    wp_sync();
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    MGQP_UNROLL2
    for (int i = 0; i < n; i++) {
      a0 = d_add(a0, sred[i]);
      a1 = d_add(a1, sgc[i]);
      a2 = d_add(a2, sgs[i]);
    }
    wp_sync();
    r0 = a0;
    r1 = a1;
    r2 = a2;
  };

  int status = ST_OK;
  bool done = false;
  double f = 0.0;

  // ------------------------------------------------------------------ setup (Appendix A, top)
  double c1 = 0.0;
  MGQP_UNROLL1
  for (int i = 0; i < n; i++) c1 = d_add(c1, sG[i * n + i]);
  if (in_n) {
    MGQP_UNROLL1
    for (int j = 0; j < n; j++) sR[li * ldR + j] = 0.0;
  }

  // G == I exactly (diagonal 1.0, off-diagonal +0.0, bit for bit) is what the reference's caller
  // always passes (src/mgqp.cpp:672, :700: JG = Identity).  Then the Cholesky factor is I and
  // J = L^-T = I with +0.0 zeros, exactly as the general code would produce them (every update is
  // x - (+0.0 * y) = x and every quotient is x / 1.0), so the factorisation and the n forward
  // eliminations can be skipped when EVERY QP of the warp has that form.
  bool row_is_identity = true;
  if (in_n) {
    MGQP_UNROLL1
    for (int j = 0; j < n; j++) {
      const double gv = sG[li * n + j];
      row_is_identity = row_is_identity && ((j == li) ? (gv == 1.0) : d_is_pos_zero(gv));
    }
  }
  const bool all_identity = !wp_any(!row_is_identity);

  // Cholesky, row-oriented as the reference: lane j owns entry (i, j) of row i
  MGQP_UNROLL1
  for (int i = 0; i < (all_identity ? 0 : n); i++) {
    const bool act = in_n && li >= i;
    double sum = 0.0;
    if (act) {
      sum = sG[i * n + li];
      const double* gi = sG + i * n;
      const double* gl = sG + li * n;
      MGQP_UNROLL2
      for (int k = i - 1; k >= 0; k--) sum = msub<FMA>(sum, gi[k], gl[k]);
    }
    const double piv = wp_shfl(sum, i, GW);
    if (!done && piv <= 0.0) {
      status = ST_NOT_PD;
      f = piv;
      done = true;
    }
    const double dii = d_sqrt(piv);
    if (act) {
      if (li == i) {
        sG[i * n + i] = dii;
      } else {
        const double v = d_div(sum, dii);
        sG[li * n + i] = v;
        sG[i * n + li] = v;
      }
    }
    wp_sync();
  }

  // J = L^-T: lane i runs forward_elimination(L, e_i) and keeps the result as row i of J
  if (in_n && all_identity) {
    MGQP_UNROLL1
    for (int r = 0; r < n; r++) sJ[li * ldJ + r] = (r == li) ? 1.0 : 0.0;
  } else if (in_n) {
    MGQP_UNROLL1
    for (int r = 0; r < n; r++) {
      double acc = (r == li) ? 1.0 : 0.0;
      const double* grow = sG + r * n;
      const double* jrow = sJ + li * ldJ;
      MGQP_UNROLL2
      for (int j = 0; j < r; j++) acc = msub<FMA>(acc, grow[j], jrow[j]);
      sJ[li * ldJ + r] = d_div(acc, sG[r * n + r]);
    }
  }
  wp_sync();
  double c2 = 0.0;
  MGQP_UNROLL1
  for (int i = 0; i < n; i++) c2 = d_add(c2, sJ[i * ldJ + i]);

  // x = -G^-1 g0 : forward elimination across lanes (lane i accumulates row i; the subtraction
  // order j = 0,1,.. is the order in which y_j become available), backward elimination serially
  // (its ascending-j subtraction order is the reverse of availability, so it cannot be spread).
  double xv = 0.0;
  {
    double acc = in_n ? sg0[li] : 0.0, yv = 0.0;
    MGQP_UNROLL1
    for (int j = 0; j < n; j++) {
      const double q = d_div_z(acc, sG[j * n + j]);  // g0 = 0 in the reference's own caller (src/mgqp.cpp:706)
      const double yj = wp_shfl(q, j, GW);
      if (li == j) yv = yj;
      if (in_n && li > j) acc = msub<FMA>(acc, sG[li * n + j], yj);
    }
    if (in_n) sd[li] = yv;
    wp_sync();
    MGQP_UNROLL1
    for (int i = n - 1; i >= 0; i--) {
      double a = sd[i];
      const double* grow = sG + i * n;
      MGQP_UNROLL2
      for (int j = i + 1; j < n; j++) a = msub<FMA>(a, grow[j], sx[j]);
      a = d_div_z(a, sG[i * n + i]);
      sx[i] = a;  // every lane of the group stores the same value
      if (li == i) xv = -a;
    }
    wp_sync();
    if (in_n) sx[li] = xv;
    wp_sync();
    double fs = 0.0;
    MGQP_UNROLL1
    for (int i = 0; i < n; i++) fs = mad<FMA>(sg0[i], sx[i], fs);
    if (!done) f = d_mul(0.5, fs);
  }

  // ------------------------------------------------------------------ working-set state
  int iq = 0;            // |A|
  double uv = 0.0;       // u[li]
  int Av = 0;            // A[li]
  double uq = 0.0;       // u[iq]  (dual of the constraint being added)
  int Aq = 0;            // A[iq]
  double R_norm = 1.0;
  unsigned act_mask = 0u;        // bit k: inequality li + k*GW is in A   (iai[i] == -1)
  unsigned excl_mask = ~0u;      // bit k: iaexcl[li + k*GW]

  // d = J^T np ; z = J[:, iq:] d[iq:] ; r = R[0:iq,0:iq]^-1 d[0:iq]
  double dv = 0.0, zv = 0.0, rv = 0.0, npv = 0.0;
  auto direction = [&](const double* npp, int nps) {
    dv = 0.0;
    npv = in_n ? npp[li * nps] : 0.0;
    if (in_n) {
      const double* pj = sJ + li;
      const double* pn = npp;
      MGQP_UNROLL_HOT
      for (int j = 0; j < n; j++, pj += ldJ, pn += nps) dv = mad<FMA>(*pj, *pn, dv);
    }
    if (in_n) sd[li] = dv;
#if MGQP_SMALL_SPLIT_DIV
    // refined reciprocals of R's diagonal for the back-substitution below, one per lane (the coefficient row sgx
    // is free outside add_constraint)
    if (li < iq) sgx[li] = d_rcp_refined(sR[li * ldR + li]);
#endif
    wp_sync();
    zv = 0.0;
    if (in_n) {
      const double* pj = sJ + li * ldJ;
      MGQP_UNROLL_HOT
      for (int j = iq; j < n; j++) zv = mad<FMA>(pj[j], sd[j], zv);
    }
    rv = 0.0;
    MGQP_UNROLL1
    for (int i = iq - 1; i >= 0; i--) {
      double sum = 0.0;
      const double* ri_row = sR + i * ldR;
      MGQP_UNROLL2
      for (int j = i + 1; j < iq; j++) sum = mad<FMA>(ri_row[j], sr[j], sum);
#if MGQP_SMALL_SPLIT_DIV
      // This loop is one dependency chain that every lane of the group repeats: the quotient on the split inline
      // path of the division (warp_prims.cuh) is three dependent operations instead of a dozen, the refinement of
      // 1 / r_ii having been done off the chain.  Outside that path's range (an exact zero, as sparse problems
      // produce them all the time) the plain division takes over.
      const double num = d_sub(sd[i], sum), den = sR[i * ldR + i];
      bool okq = true;
      double ri = d_div_rcp(num, den, sgx[i], okq);
      if (!okq) ri = d_div_z(num, den);
#else
      const double ri = d_div_z(d_sub(sd[i], sum), sR[i * ldR + i]);
#endif
      sr[i] = ri;  // same value from every lane of the group
      if (li == i) rv = ri;
    }
  };

  // add_constraint for the groups with `doit`.  Returns the reference's bool.
  // Lock step over j = n-1 .. min(iq)+1; a group takes part while j >= its own iq+1.
  auto add_constraint = [&](bool doit) -> bool {
    const int jlo = doit ? iq + 1 : n;
    const int jmin = wp_min(jlo);
    // (1) the serial part of the sweep is only the chain of hypotenuses: `distance` uses |a|, |b|,
    // so h_j = distance(d[j-1], h_{j+1}) needs neither the signs nor the divisions of rotation j+1.
    // Every lane of the group runs the chain; lane j keeps h_j.
    double carry = (n > 0) ? d_abs(sd[n - 1]) : 0.0;
    double hreg = 1.0;
    unsigned skipm = 0u;
    MGQP_UNROLL1
    for (int j = n - 1; j >= jmin; j--) {
      if (j >= jlo) {
        const double h = gi_distance(sd[j - 1], carry);
        if (d_abs(h) < MGQP_EPS) {
          skipm |= 1u << j;
          carry = d_abs(sd[j - 1]);
        } else {
          if (li == j) hreg = h;
          carry = h;
        }
      }
    }
    // (2) lane j turns (d[j-1], incoming d[j], h_j) into the rotation's (c, s, s/(1+c)) with the
    // reference's sign normalisation; the incoming d[j] is what rotation j+1 leaves behind.
    const bool myrot = li >= jlo && li <= n - 1 && !((skipm >> li) & 1u);
    const double dprev = (li > 0 && in_n) ? sd[li - 1] : 0.0;
#if MGQP_SMALL_SPLIT_DIV
    const double rh = d_rcp_refined(hreg);  // shared by c and s
    bool okc = true;
    double c = d_div_rcp(dprev, hreg, rh, okc);
    if (!okc) c = d_div_z(dprev, hreg);
#else
    double c = d_div_z(dprev, hreg);
#endif
    const double dnew = ((skipm >> li) & 1u) ? dprev : ((c < 0.0) ? -hreg : hreg);  // new d[li-1]
    const double dnext = wp_shfl_down(dnew, 1, GW);
    const double din = (li == n - 1) ? sd[n - 1] : dnext;
#if MGQP_SMALL_SPLIT_DIV
    bool oks = true;
    double sn = d_div_rcp(din, hreg, rh, oks);
    if (!oks) sn = d_div_z(din, hreg);
#else
    double sn = d_div_z(din, hreg);
#endif
    if (c < 0.0) {
      c = -c;
      sn = -sn;
    }
    if (myrot) {
      sgc[li] = c;
      sgs[li] = sn;
#if MGQP_SMALL_SPLIT_DIV
      const double den = d_add(1.0, c);
      bool okx = true;
      double xq = d_div_rcp(sn, den, d_rcp_refined(den), okx);
      if (!okx) xq = d_div_z(sn, den);
      sgx[li] = xq;
#else
      sgx[li] = d_div_z(sn, d_add(1.0, c));
#endif
    }
    const double dl_rot = wp_shfl(dnew, jlo & (GW - 1), GW);
    const double dlast = (jlo <= n - 1) ? dl_rot : sd[iq < n ? iq : 0];
    wp_sync();
    // (3) lane k streams row k of J right to left through all rotations; no synchronisation needed
    if (in_n) {
      double* jrow = sJ + li * ldJ;
      double t2 = jrow[n - 1];
      MGQP_UNROLL_HOT
      for (int j = n - 1; j >= jlo; j--) {
        const double t1 = jrow[j - 1];
        if (!((skipm >> j) & 1u)) {
          const double n1 = d_add(d_mul(t1, sgc[j]), d_mul(t2, sgs[j]));
          jrow[j] = d_sub(d_mul(sgx[j], d_add(t1, n1)), t2);
          t2 = n1;
        } else {
          jrow[j] = t2;
          t2 = t1;
        }
      }
      if (jlo <= n - 1) jrow[jlo - 1] = t2;
    }
    bool ok = true;
    if (doit) {
      // column iq of R <- d[0..iq]   (d[iq] is what the last rotation left)
      if (li < iq) sR[li * ldR + iq] = dv;
      if (li == iq) sR[li * ldR + iq] = dlast;
      iq++;
      if (d_abs(dlast) <= d_mul(MGQP_EPS, R_norm)) {
        ok = false;
      } else {
        R_norm = (R_norm < d_abs(dlast)) ? d_abs(dlast) : R_norm;
      }
    }
    wp_sync();
    return ok;
  };

  // delete_constraint(l) for the groups with `doit`.
  auto delete_constraint = [&](bool doit, int l) {
    const unsigned bal = wp_ballot(doit && li >= p && li < iq && Av == l);
    const unsigned bits = (GW == 32) ? bal : ((bal >> (g * GW)) & ((1u << (GW & 31)) - 1u));
    const int qq = bits ? (wp_ffs(bits) - 1) : 0;
    const int A_dn = wp_shfl_down(Av, 1, GW);
    const double u_dn = wp_shfl_down(uv, 1, GW);
    if (doit) {
      if (li >= qq && li < iq - 1) {
        Av = A_dn;
        uv = u_dn;
      }
      // The reference moves the pending pair (A[iq], u[iq]) down one slot; here the pending pair
      // lives in (Aq, uq) = "slot iq" for whatever iq currently is, so it simply stays put.
      if (li == iq - 1) {
        Av = Aq;
        uv = uq;
      }
      if (in_n) {
        MGQP_UNROLL1
        for (int i = qq; i < iq - 1; i++) sR[li * ldR + i] = sR[li * ldR + i + 1];
        if (li < iq) sR[li * ldR + iq - 1] = 0.0;
      }
      iq--;
    }
    const int len = (doit && iq > 0) ? (iq - qq) : 0;
    const int lmax = wp_max(len);
    wp_sync();
    MGQP_UNROLL1
    for (int t = 0; t < lmax; t++) {
      const int j = qq + t;
      bool rot = false;
      double cc = 0.0, ss = 0.0, xny = 0.0, diag = 0.0;
      if (t < len) {
        cc = sR[j * ldR + j];
        ss = sR[(j + 1) * ldR + j];
        rot = gi_givens(cc, ss, xny, diag);
      }
      wp_sync();  // pivot pair read by everyone before lane j / j+1 overwrite it
      if (rot) {
        if (li > j && li < iq) {
          const double t1 = sR[j * ldR + li], t2 = sR[(j + 1) * ldR + li];
          const double n1 = d_add(d_mul(t1, cc), d_mul(t2, ss));
          sR[j * ldR + li] = n1;
          sR[(j + 1) * ldR + li] = d_sub(d_mul(xny, d_add(t1, n1)), t2);
        }
        if (li == j) {
          sR[(j + 1) * ldR + j] = 0.0;
          sR[j * ldR + j] = diag;
        }
        if (in_n) {
          const double t1 = sJ[li * ldJ + j], t2 = sJ[li * ldJ + j + 1];
          const double n1 = d_add(d_mul(t1, cc), d_mul(t2, ss));
          sJ[li * ldJ + j] = n1;
          sJ[li * ldJ + j + 1] = d_sub(d_mul(xny, d_add(n1, t1)), t2);
        }
      }
      wp_sync();
    }
  };

  // ------------------------------------------------------------------ equality constraints
  MGQP_UNROLL1
  for (int e = 0; e < p; e++) {
    direction(sCE + e, p);
    double zz, npx, znp;
    osum3(d_mul(zv, zv), d_mul(npv, xv), d_mul(zv, npv), zz, npx, znp);
    double t2 = 0.0;
    if (d_abs(zz) > MGQP_EPS) t2 = d_div(d_sub(-npx, sce0[e]), znp);
    if (!done) {
      xv = d_add(xv, d_mul(t2, zv));
      uq = t2;
      if (li < iq) uv = d_sub(uv, d_mul(t2, rv));
      f = d_add(f, d_mul(d_mul(0.5, d_mul(t2, t2)), znp));
      Aq = -e - 1;
    }
    const int iq_before = iq;
    const bool ok = add_constraint(!done);
    if (!done) {
      if (li == iq_before) {
        uv = uq;
        Av = Aq;
      }
      if (!ok) {
        status = ST_EQ_DEPENDENT;
        done = true;
      }
    }
  }
  if (in_n) sx[li] = xv;
  wp_sync();
  // the state a retry without inequalities would return (see TileOut::fuse_retry)
  const bool fused = out.fuse_retry && all_identity;
  const double x_eq = xv, f_eq = f;
  const int iq_eq = iq, A_eq = Av;
  const bool done_eq = done;

  // ------------------------------------------------------------------ active-set iteration
  const double thr = d_mul(d_mul(d_mul(d_mul((double)m, MGQP_EPS), c1), c2), 100.0);
  const int nk = (m + GW - 1) >> LOGG;
  int ip = 0, steps = 0;
  double ss = 0.0, s_ip = 0.0;
  bool need_scan = true, l1_mode = true;
  double x_old = 0.0, u_old = 0.0;
  int A_old = 0;

  MGQP_UNROLL1
  while (wp_any(!done)) {
    // ---------------- step 1 / step 2: violated-constraint scan and choice of ip
    if (wp_any(!done && need_scan)) {
      const bool scan = !done && need_scan;
      if (scan && l1_mode) {
        excl_mask = ~0u;
        ss = 0.0;
        ip = 0;
      }
      double best = ss, psi = 0.0;
      int bidx = NONE;
      int nbad = 0, kbad = 0;
      if (comp) {
        const unsigned bal = wp_ballot(in_n && !(d_abs(xv) < INF));
        const unsigned bits = (GW == 32) ? bal : ((bal >> (g * GW)) & ((1u << (GW & 31)) - 1u));
        nbad = i_popc(bits);
        kbad = bits ? (wp_ffs(bits) - 1) : 0;
      }
      MGQP_UNROLL1
      for (int k = 0; k < nk; k++) {
        const int i = li + (k << LOGG);
        if (i < m) {
          double sum = 0.0;
          if (comp) {
            sum = comp_dot(i, nbad, kbad);
          } else {
            const double* pc = sCI + i;
            MGQP_UNROLL_HOT
            for (int j = 0; j < n; j++, pc += m) sum = mad<FMA>(*pc, sx[j], sum);
          }
          sum = d_add(sum, sci0[i]);
          if (sum < 0.0) psi = d_add(psi, sum);
          if (sum < best && !((act_mask >> k) & 1u) && ((excl_mask >> k) & 1u)) {
            best = sum;
            bidx = i;
          }
        }
      }
      MGQP_UNROLL1
      for (int off = GW >> 1; off > 0; off >>= 1) {
        const double ov = wp_shfl_xor(best, off, GW);
        const int oi = wp_shfl_xor(bidx, off, GW);
        psi = d_add(psi, wp_shfl_xor(psi, off, GW));
        if (ov < best || (ov == best && oi < bidx)) {
          best = ov;
          bidx = oi;
        }
      }
      if (scan) {
        if (l1_mode) {
          if (!FMA && d_abs(psi) > d_mul(0.5, thr) && d_abs(psi) < d_mul(2.0, thr)) {
            // within 2x of the threshold: redo the sum in the reference's order (rare)
            psi = 0.0;
            MGQP_UNROLL1
            for (int i = 0; i < m; i++) {
              double sum = 0.0;
              if (comp) {
                sum = comp_dot(i, nbad, kbad);
              } else {
                MGQP_UNROLL1
                for (int j = 0; j < n; j++) sum = mad<FMA>(sCI[j * m + i], sx[j], sum);
              }
              sum = d_add(sum, sci0[i]);
              psi = d_add(psi, (sum < 0.0) ? sum : 0.0);
            }
          }
          if (d_abs(psi) <= thr) {
            done = true;  // numerically no infeasibility left
          } else {
            u_old = uv;
            A_old = Av;
            x_old = xv;
          }
        }
        if (!done) {
          if (bidx != NONE) {
            ss = best;
            ip = bidx;
            s_ip = best;
          }
          if (ss >= 0.0) {
            done = true;
          } else {
            uq = 0.0;
            Aq = ip;
            need_scan = false;
          }
        }
      }
    }

    // ---------------- step 2a: step direction
    if (!done) {
      steps++;
      if (steps > max_iter) {
        status = ST_ITER_CAP;
        done = true;
      }
    }
    if (!wp_any(!done)) break;
    if (comp) {
      // the pending column as a vector, in r's place (free until update_r rewrites it at the end of direction)
      if (in_n) sr[li] = (m > 0 && li == sCIr[ip]) ? sCI[ip] : 0.0;
      wp_sync();
      direction(sr, 1);
    } else {
      direction(sCI + ip, m);
    }

    // ---------------- step 2b: step lengths
    double ratio = INF;
    if (li >= p && li < iq && rv > 0.0) {
      ratio = d_div(uv, rv);
      if (!(ratio < INF)) ratio = INF;  // NaN / inf never win `u/r < t1`
    }
    int lk = (ratio < INF) ? li : NONE;
    MGQP_UNROLL1
    for (int off = GW >> 1; off > 0; off >>= 1) {
      const double ov = wp_shfl_xor(ratio, off, GW);
      const int oi = wp_shfl_xor(lk, off, GW);
      if (ov < ratio || (ov == ratio && oi < lk)) {
        ratio = ov;
        lk = oi;
      }
    }
    const double t1 = ratio;
    const int l_src = wp_shfl(Av, (lk == NONE) ? 0 : lk, GW);
    const int l = (lk == NONE) ? 0 : l_src;
    double zz, znp;
    osum2(d_mul(zv, zv), d_mul(zv, npv), zz, znp);
    double t2;
    if (d_abs(zz) > MGQP_EPS) {
      t2 = d_div(-s_ip, znp);
      if (t2 < 0.0) t2 = INF;
    } else {
      t2 = INF;
    }
    const double t = (t2 < t1) ? t2 : t1;  // std::min(t1, t2)

    // ---------------- step 2c: take the step
    bool do_add = false, do_del = false, partial = false;
    int del_l = 0;
    if (!done) {
      if (t >= INF) {
        status = ST_INFEASIBLE;
        f = INF;
        done = true;
      } else if (t2 >= INF) {
        // dual step only: drop constraint l
        if (li < iq) uv = d_sub(uv, d_mul(t, rv));
        uq = d_add(uq, t);
        do_del = true;
        del_l = l;
      } else {
        xv = d_add(xv, d_mul(t, zv));
        f = d_add(f, d_mul(d_mul(t, znp), d_add(d_mul(0.5, t), uq)));
        if (li < iq) uv = d_sub(uv, d_mul(t, rv));
        uq = d_add(uq, t);
        if (d_abs(d_sub(t, t2)) < MGQP_EPS) {
          do_add = true;  // full step
        } else {
          do_del = true;  // partial step: drop l, stay on ip
          partial = true;
          del_l = l;
        }
      }
    }
    if (do_del) {  // iai[l] = l
      if (li == (del_l & (GW - 1))) act_mask &= ~(1u << (del_l >> LOGG));
    }

    bool degenerate = false;
    if (wp_any(do_add)) {
      const int iq_before = iq;
      const bool ok = add_constraint(do_add);
      if (do_add) {
        if (li == iq_before) {
          uv = uq;
          Av = Aq;
        }
        if (ok) {
          if (li == (ip & (GW - 1))) act_mask |= (1u << (ip >> LOGG));  // iai[ip] = -1
          need_scan = true;
          l1_mode = true;
        } else {
          degenerate = true;
          if (li == (ip & (GW - 1))) excl_mask &= ~(1u << (ip >> LOGG));  // iaexcl[ip] = false
          do_del = true;
          del_l = ip;
        }
      }
    }
    if (wp_any(do_del)) delete_constraint(do_del, del_l);

    if (wp_any(degenerate)) {
      // restore the state saved at step 1 and go back to step 2 with the stale ss
      if (degenerate) {
        act_mask = 0u;
        if (li >= p && li < iq) {
          Av = A_old;
          uv = u_old;
        }
        xv = x_old;
      }
      MGQP_UNROLL1
      for (int k = 0; k < n; k++) {  // iai[A[i]] = -1 for i in [p, iq)
        const int a = wp_shfl(Av, k, GW);
        if (degenerate && k >= p && k < iq && li == (a & (GW - 1))) act_mask |= (1u << (a >> LOGG));
      }
      if (degenerate) {
        need_scan = true;
        l1_mode = false;
      }
    }
    if (in_n) sx[li] = xv;
    if (wp_any(partial)) {
      // s[ip] = CI[:,ip].x + ci0[ip] with the new x
      const double cip = comp ? ((in_n && li == sCIr[ip]) ? sCI[ip] : 0.0) : (in_n ? sCI[li * m + ip] : 0.0);
      const double v = osum(d_mul(cip, xv));
      if (partial) s_ip = d_add(v, sci0[ip]);
    }
    wp_sync();
  }

  // ------------------------------------------------------------------ results
  if (fused && !done_eq && (f != f || f == d_inf()) && status != ST_NOT_PD && status != ST_EQ_DEPENDENT &&
      f_eq == f_eq && d_abs(f_eq) < d_inf()) {
    // (a retry whose own cost is NaN / inf is left to the generic pass: it zeroes the command)
    xv = x_eq;
    f = f_eq;
    iq = iq_eq;
    Av = A_eq;
    steps = 0;
    status = 16;  // MGQP_STATUS_RETRIED | ok
  }
  if (g < cnt) {
    if (in_n) out.x[g * n + li] = xv;
    if (li == 0) {
      out.f[g] = f;
      out.status[g] = status;
      if (out.iq) out.iq[g] = iq;
      if (out.iters) out.iters[g] = steps;
    }
    if (out.A) {
      int* Ag = out.A + (long long)g * (p + m);
      MGQP_UNROLL1
      for (int i = li; i < p + m; i += GW) Ag[i] = (i < iq) ? Av : NONE;  // i < iq <= n implies i == li
    }
  }
}

}  // namespace mgqp
