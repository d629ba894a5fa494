// Goldfarb-Idnani dual active-set solver, ONE QP PER LANE (n <= 8, m <= 64): a warp advances 32
// independent QPs in SIMT fashion.
//
// Replaces the body of the reference's solve_quadprog (lib/QuadProgpp/libquadprog.a: QuadProg++.o,
// prototype include/QuadProgpp/QuadProg++.hh:69-72, behavioural spec SURVEY.md Appendix A) for the
// small joint-space shapes (BASELINE.json configs[1]/[4]: n=7, p=1, m=42).
//
// Why a second kernel family: with n = 7 a lane group of 8 (gi_small.cuh) spends most of its issue
// slots on work that is uniform across the group (divisions, square roots, control) and on ordered
// cross-lane sums; one lane per QP executes each of those once per QP, about 7x fewer warp
// instructions per QP (profiles/r01_notes.md).  The price is shared memory: a lane owns a private,
// bank-skewed slot, so the constraint matrix cannot stay resident in dense form.  Instead each
// inequality column is classified while it streams in from HBM (its only read):
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
// Slot layout (doubles, lane stride == 1 mod 16 so that the 32 lanes of a warp hit 32 distinct
// bank pairs whenever they address the same element):
//   J[N*N] | R[N*N] (G lands here, is factored in place, then becomes R) | x[N] | d[N] | r[N] |
//   u[N+1] | u_old[N+1] | civ[m] | ci0[m] | ints: A[N+1], A_old[N+1], cidx[m]
#pragma once
#include "warp_prims.cuh"

#ifndef MGQP_EPS
#define MGQP_EPS 2.220446049250313e-16
#endif

#ifdef MGQP_EMULATE
#define MGQP_LDG(p) (*(p))
#else
#define MGQP_LDG(p) __ldg(p)
#endif

namespace mgqp {

struct LaneView {
  int n, p, m;
  int oJ, oR, ox, od, orr, ou, ouold, ociv, oci0, oInt;  // doubles
  int iA, iAold, icidx;                                   // ints, relative to (int*)(slot + oInt)
  int stride;                                             // doubles per lane, == 1 (mod 16)
};

MGQP_HOST inline LaneView make_lane_view(int n, int p, int m) {
  LaneView v;
  v.n = n; v.p = p; v.m = m;
  int o = 0;
  v.oJ = o;     o += n * n;
  v.oR = o;     o += n * n;
  v.ox = o;     o += n;
  v.od = o;     o += n;
  v.orr = o;    o += n;
  v.ou = o;     o += n + 1;
  v.ouold = o;  o += n + 1;
  v.ociv = o;   o += m;
  v.oci0 = o;   o += m;
  v.oInt = o;
  v.iA = 0;
  v.iAold = n + 1;
  v.icidx = 2 * (n + 1);
  const int ints = 2 * (n + 1) + m;
  o += (ints + 1) / 2;
  // staging needs room for ceil(n*m/3) raw doubles in J|R
  while ((o & 15) != 1) o++;
  v.stride = o;
  return v;
}

enum { LN_EMPTY = -2, LN_DENSE = -1 };

MGQP_DEV double ln_distance(double a, double b) {
  const double a1 = d_abs(a), b1 = d_abs(b);
  if (a1 > b1) {
    const double t = d_div(b1, a1);
    return d_mul(a1, d_sqrt(d_add(1.0, d_mul(t, t))));
  } else if (b1 > a1) {
    const double t = d_div(a1, b1);
    return d_mul(b1, d_sqrt(d_add(1.0, d_mul(t, t))));
  }
  return d_mul(a1, d_sqrt(2.0));
}

struct LaneIO {  // global memory, this lane's QP
  const double* G; const double* g0; const double* CE; const double* ce0; const double* CI; const double* ci0;
};

// Classify `cnt` consecutive raw elements of CI (row-major element index e0 .. e0+cnt) that were
// staged at S[stage..]: first nonzero of a column -> (row, value); a second one -> DENSE.
MGQP_DEV void ln_classify(double* S, const LaneView& v, int stage, int e0, int cnt) {
  int* I = reinterpret_cast<int*>(S + v.oInt);
  int j = e0 / v.m, i = e0 - j * v.m;
#pragma unroll 1
  for (int k = 0; k < cnt; k++) {
    const double a = S[stage + k];
    if (a != 0.0) {  // NaN counts as nonzero
      const int c = I[v.icidx + i];
      if (c == LN_EMPTY) {
        I[v.icidx + i] = j;
        S[v.ociv + i] = a;
      } else {
        I[v.icidx + i] = LN_DENSE;
      }
    }
    if (++i == v.m) { i = 0; j++; }
  }
}

// The solver proper.  Preconditions: S[oR..] holds G (row-major n*n), S[od..] holds g0,
// S[oci0..] holds ci0, civ/cidx hold the classified columns.  N == v.n.
template <int N, bool FMA>
MGQP_DEV void gi_lane_solve(const LaneView& v, double* S, const LaneIO& io, int max_iter, double* x_out,
                            double& f_out, int& status_out, int& iq_out, int& iters_out) {
  const int p = v.p, m = v.m;
  double* J = S + v.oJ;
  double* R = S + v.oR;  // L during setup
  double* sx = S + v.ox;
  double* sd = S + v.od;
  double* sr = S + v.orr;
  double* su = S + v.ou;
  double* suold = S + v.ouold;
  const double* civ = S + v.ociv;
  const double* sci0 = S + v.oci0;
  int* I = reinterpret_cast<int*>(S + v.oInt);
  int* A = I + v.iA;
  int* Aold = I + v.iAold;
  const int* cidx = I + v.icidx;
  const double INF = d_inf();

  int status = 0, iq = 0, steps = 0;
  double f = 0.0;
  double x[N], z[N], np[N], dd[N], xold[N];
#pragma unroll
  for (int i = 0; i < N; i++) { x[i] = 0.0; z[i] = 0.0; np[i] = 0.0; dd[i] = 0.0; xold[i] = 0.0; }

  // ---------------------------------------------------------------- setup
  double c1 = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) c1 = d_add(c1, R[i * N + i]);

  bool fail = false;
#pragma unroll 1
  for (int i = 0; i < N && !fail; i++) {
#pragma unroll 1
    for (int j = i; j < N; j++) {
      double sum = R[i * N + j];
#pragma unroll 1
      for (int k = i - 1; k >= 0; k--) sum = msub<FMA>(sum, R[i * N + k], R[j * N + k]);
      if (i == j) {
        if (sum <= 0.0) {
          status = 2;
          f = sum;
          fail = true;
          break;
        }
        R[i * N + i] = d_sqrt(sum);
      } else {
        R[j * N + i] = d_div(sum, R[i * N + i]);
      }
    }
    if (!fail) {
#pragma unroll 1
      for (int k = i + 1; k < N; k++) R[i * N + k] = R[k * N + i];
    }
  }

  if (!fail) {
    // J = L^-T (row i = forward_elimination(L, e_i)); c2 = trace(J)
    double c2 = 0.0;
#pragma unroll 1
    for (int i = 0; i < N; i++) {
#pragma unroll 1
      for (int r = 0; r < N; r++) {
        double acc = (r == i) ? 1.0 : 0.0;
#pragma unroll 1
        for (int j = 0; j < r; j++) acc = msub<FMA>(acc, R[r * N + j], J[i * N + j]);
        J[i * N + r] = d_div(acc, R[r * N + r]);
      }
      c2 = d_add(c2, J[i * N + i]);
    }
    // x = -G^-1 g0 ; sd holds g0 on entry, y is built in sr, x in sx
    sr[0] = d_div(sd[0], R[0]);
#pragma unroll 1
    for (int i = 1; i < N; i++) {
      double a = sd[i];
#pragma unroll 1
      for (int j = 0; j < i; j++) a = msub<FMA>(a, R[i * N + j], sr[j]);
      sr[i] = d_div(a, R[i * N + i]);
    }
    sx[N - 1] = d_div(sr[N - 1], R[(N - 1) * N + (N - 1)]);
#pragma unroll 1
    for (int i = N - 2; i >= 0; i--) {
      double a = sr[i];
#pragma unroll 1
      for (int j = i + 1; j < N; j++) a = msub<FMA>(a, R[i * N + j], sx[j]);
      sx[i] = d_div(a, R[i * N + i]);
    }
    double fs = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      x[i] = -sx[i];
      fs = mad<FMA>(sd[i], x[i], fs);
    }
    f = d_mul(0.5, fs);
    // L is dead: the block becomes R (all zero)
#pragma unroll 1
    for (int i = 0; i < N * N; i++) R[i] = 0.0;

    double R_norm = 1.0;

    // d = J^T np (registers + smem copy), z = J[:, iq:] d[iq:], r = R^-1 d
    auto direction = [&]() {
#pragma unroll
      for (int i = 0; i < N; i++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < N; j++) s = mad<FMA>(J[j * N + i], np[j], s);
        dd[i] = s;
        sd[i] = s;
      }
#pragma unroll
      for (int i = 0; i < N; i++) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < N; j++)
          if (j >= iq) s = mad<FMA>(J[i * N + j], dd[j], s);
        z[i] = s;
      }
#pragma unroll 1
      for (int i = iq - 1; i >= 0; i--) {
        double s = 0.0;
#pragma unroll 1
        for (int j = i + 1; j < iq; j++) s = mad<FMA>(R[i * N + j], sr[j], s);
        sr[i] = d_div(d_sub(sd[i], s), R[i * N + i]);
      }
    };
    auto dot_zz = [&]() {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < N; i++) s = mad<FMA>(z[i], z[i], s);
      return s;
    };
    auto dot_znp = [&]() {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < N; i++) s = mad<FMA>(z[i], np[i], s);
      return s;
    };
    auto dot_npx = [&]() {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < N; i++) s = mad<FMA>(np[i], x[i], s);
      return s;
    };

    // add_constraint (Appendix A): Givens sweep j = n-1 .. iq+1 on d, columns j-1, j of J
    auto add_constraint = [&]() -> bool {
#pragma unroll 1
      for (int j = N - 1; j >= iq + 1; j--) {
        double cc = sd[j - 1], ss = sd[j];
        const double h = ln_distance(cc, ss);
        if (d_abs(h) < MGQP_EPS) continue;
        sd[j] = 0.0;
        ss = d_div(ss, h);
        cc = d_div(cc, h);
        if (cc < 0.0) {
          cc = -cc;
          ss = -ss;
          sd[j - 1] = -h;
        } else {
          sd[j - 1] = h;
        }
        const double xny = d_div(ss, d_add(1.0, cc));
#pragma unroll
        for (int k = 0; k < N; k++) {
          const double t1 = J[k * N + j - 1], t2 = J[k * N + j];
          const double n1 = d_add(d_mul(t1, cc), d_mul(t2, ss));
          J[k * N + j - 1] = n1;
          J[k * N + j] = d_sub(d_mul(xny, d_add(t1, n1)), t2);
        }
      }
      iq++;
#pragma unroll 1
      for (int i = 0; i < iq; i++) R[i * N + iq - 1] = sd[i];
      const double dl = d_abs(sd[iq - 1]);
      if (dl <= d_mul(MGQP_EPS, R_norm)) return false;
      R_norm = (R_norm < dl) ? dl : R_norm;
      return true;
    };

    // delete_constraint(l) (Appendix A).  A/u have a valid slot at index iq (the pending pair).
    auto delete_constraint = [&](int l) {
      int qq = 0;
#pragma unroll 1
      for (int i = p; i < iq; i++)
        if (A[i] == l) {
          qq = i;
          break;
        }
#pragma unroll 1
      for (int i = qq; i < iq - 1; i++) {
        A[i] = A[i + 1];
        su[i] = su[i + 1];
#pragma unroll
        for (int j = 0; j < N; j++) R[j * N + i] = R[j * N + i + 1];
      }
      A[iq - 1] = A[iq];
      su[iq - 1] = su[iq];
      A[iq] = 0;
      su[iq] = 0.0;
#pragma unroll 1
      for (int j = 0; j < iq; j++) R[j * N + iq - 1] = 0.0;
      iq--;
      if (iq == 0) return;
#pragma unroll 1
      for (int j = qq; j < iq; j++) {
        double cc = R[j * N + j], ss = R[(j + 1) * N + j];
        const double h = ln_distance(cc, ss);
        if (d_abs(h) < MGQP_EPS) continue;
        cc = d_div(cc, h);
        ss = d_div(ss, h);
        R[(j + 1) * N + j] = 0.0;
        if (cc < 0.0) {
          R[j * N + j] = -h;
          cc = -cc;
          ss = -ss;
        } else {
          R[j * N + j] = h;
        }
        const double xny = d_div(ss, d_add(1.0, cc));
#pragma unroll 1
        for (int k = j + 1; k < iq; k++) {
          const double t1 = R[j * N + k], t2 = R[(j + 1) * N + k];
          const double n1 = d_add(d_mul(t1, cc), d_mul(t2, ss));
          R[j * N + k] = n1;
          R[(j + 1) * N + k] = d_sub(d_mul(xny, d_add(t1, n1)), t2);
        }
#pragma unroll
        for (int k = 0; k < N; k++) {
          const double t1 = J[k * N + j], t2 = J[k * N + j + 1];
          const double n1 = d_add(d_mul(t1, cc), d_mul(t2, ss));
          J[k * N + j] = n1;
          J[k * N + j + 1] = d_sub(d_mul(xny, d_add(n1, t1)), t2);
        }
      }
    };

    // ---------------------------------------------------------------- equality constraints
    bool dead = false;
#pragma unroll 1
    for (int e = 0; e < p && !dead; e++) {
#pragma unroll
      for (int j = 0; j < N; j++) np[j] = MGQP_LDG(io.CE + j * p + e);
      direction();
      double t2 = 0.0;
      if (d_abs(dot_zz()) > MGQP_EPS) t2 = d_div(d_sub(-dot_npx(), MGQP_LDG(io.ce0 + e)), dot_znp());
#pragma unroll
      for (int k = 0; k < N; k++) x[k] = d_add(x[k], d_mul(t2, z[k]));
      su[iq] = t2;
#pragma unroll 1
      for (int k = 0; k < iq; k++) su[k] = d_sub(su[k], d_mul(t2, sr[k]));
      f = d_add(f, d_mul(d_mul(0.5, d_mul(t2, t2)), dot_znp()));
      A[e] = -e - 1;
      if (!add_constraint()) {
        status = 3;
        dead = true;
      }
    }

    // ---------------------------------------------------------------- active-set iteration
    if (!dead) {
      const double thr = d_mul(d_mul(d_mul(d_mul((double)m, MGQP_EPS), c1), c2), 100.0);
      unsigned long long act = 0ull, excl = ~0ull;  // iai[i] == -1  /  iaexcl[i]
      int ip = 0;
      double ss = 0.0, s_ip = 0.0;
      bool need_scan = true, l1 = true, done = false;
#pragma unroll 1
      while (!done) {
        if (need_scan) {
          if (l1) {
            excl = ~0ull;
            ss = 0.0;
            ip = 0;
          }
          bool finite = true;
#pragma unroll
          for (int j = 0; j < N; j++) {
            sx[j] = x[j];
            finite = finite && (d_abs(x[j]) < INF);
          }
          double psi = 0.0;
#pragma unroll 1
          for (int i = 0; i < m; i++) {
            const int c = cidx[i];
            double sum;
            if (c >= 0 && finite) {
              sum = d_add(0.0, d_mul(civ[i], sx[c]));
            } else if (c == LN_EMPTY && finite) {
              sum = 0.0;
            } else {
              sum = 0.0;
#pragma unroll
              for (int j = 0; j < N; j++) sum = mad<FMA>(MGQP_LDG(io.CI + j * m + i), x[j], sum);
            }
            sum = d_add(sum, sci0[i]);
            psi = d_add(psi, (sum < 0.0) ? sum : 0.0);
            if (sum < ss && !((act >> i) & 1ull) && ((excl >> i) & 1ull)) {
              ss = sum;
              ip = i;
              s_ip = sum;
            }
          }
          if (l1) {
            if (d_abs(psi) <= thr) {
              done = true;
            } else {
#pragma unroll 1
              for (int i = 0; i < iq; i++) {
                suold[i] = su[i];
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
              const int c = cidx[ip];
#pragma unroll
              for (int j = 0; j < N; j++) {
                if (c >= 0) np[j] = (j == c) ? civ[ip] : 0.0;
                else if (c == LN_EMPTY) np[j] = 0.0;
                else np[j] = MGQP_LDG(io.CI + j * m + ip);
              }
              su[iq] = 0.0;
              A[iq] = ip;
              need_scan = false;
            }
          }
        }
        if (done) break;

        steps++;
        if (steps > max_iter) {
          status = 4;
          break;
        }
        direction();
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
        const double znp = dot_znp();
        if (d_abs(dot_zz()) > MGQP_EPS) {
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
#pragma unroll 1
          for (int k = 0; k < iq; k++) su[k] = d_sub(su[k], d_mul(t, sr[k]));
          su[iq] = d_add(su[iq], t);
          act &= ~(1ull << l);
          delete_constraint(l);
          continue;
        }
#pragma unroll
        for (int k = 0; k < N; k++) x[k] = d_add(x[k], d_mul(t, z[k]));
        f = d_add(f, d_mul(d_mul(t, znp), d_add(d_mul(0.5, t), su[iq])));
#pragma unroll 1
        for (int k = 0; k < iq; k++) su[k] = d_sub(su[k], d_mul(t, sr[k]));
        su[iq] = d_add(su[iq], t);
        if (d_abs(d_sub(t, t2)) < MGQP_EPS) {
          if (!add_constraint()) {
            excl &= ~(1ull << ip);
            delete_constraint(ip);
            act = 0ull;
#pragma unroll 1
            for (int i = p; i < iq; i++) {
              A[i] = Aold[i];
              su[i] = suold[i];
              act |= 1ull << A[i];
            }
#pragma unroll
            for (int i = 0; i < N; i++) x[i] = xold[i];
            need_scan = true;
            l1 = false;
          } else {
            act |= 1ull << ip;
            need_scan = true;
            l1 = true;
          }
          continue;
        }
        act &= ~(1ull << l);
        delete_constraint(l);
        s_ip = d_add(dot_npx(), sci0[ip]);  // s[ip] = CI[:,ip].x + ci0[ip] with the new x
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
