/* TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, fp64, no FMA: build with -ffp-contract=off) of the Goldfarb-Idnani
 * dual active-set solver that the reference links as a prebuilt binary
 * (/root/reference/lib/QuadProgpp/libquadprog.a, member QuadProg++.o; prototype
 * include/QuadProgpp/QuadProg++.hh:69-72; the only call sites are src/mgqp.cpp:708 and :725).
 * The solver's source is NOT in the reference tree; the algorithm is QuadProg++ (L. Di Gaspero,
 * MIT, "$Id: QuadProg++.hh 232 2007-06-21", locally modified -- containers in namespace ArrayHH),
 * restated from its published form and from SURVEY.md Appendix A (which was checked against the
 * disassembly).  Helper loops that do exist as source in the tree are cited per function below
 * (include/QuadProgpp/eigen/ArrayHelpers.hh).
 *
 * PARITY PIN: tests/test_oracle_vs_ref.py runs this file against the shipped binary itself
 * (oracle/_ref/libquadprog_ref.so, built by oracle/Makefile from the files under /root/reference)
 * and requires BIT-IDENTICAL x and f plus identical verdicts on seeded batches of every
 * benchmark shape and on the known-answer cases K1..K7 of SURVEY.md section 8c; the committed
 * fixtures under tests/golden/ were produced by that binary (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * use this file.  The product (csrc/) never links or calls it.
 *
 * Beyond what the binary returns, this restatement also reports the final active set (A, iq) and
 * the step count, which is how "identical final active sets" becomes checkable.
 */
#include "gi_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* test instrumentation: how often the degenerate-add branch ran (not thread-safe; tests use 1 thread) */
long long gi_oracle_degenerate_count = 0;

#define GI_EPS 2.220446049250313e-16 /* std::numeric_limits<double>::epsilon(), .rodata+0x190 */

/* QuadProg++ `distance` (weak symbol in QuadProg++.o): overflow-safe sqrt(a^2+b^2). */
static double gi_distance(double a, double b) {
  double a1 = fabs(a), b1 = fabs(b), t;
  if (a1 > b1) {
    t = b1 / a1;
    return a1 * sqrt(1.0 + t * t);
  } else if (b1 > a1) {
    t = a1 / b1;
    return b1 * sqrt(1.0 + t * t);
  }
  return a1 * sqrt(2.0);
}

/* scalar_product: ascending index, accumulator starts at 0.0. */
static double gi_dot(const double* a, const double* b, int n) {
  double sum = 0.0;
  for (int i = 0; i < n; i++) sum += a[i] * b[i];
  return sum;
}

/* cholesky_decomposition (QuadProg++.o+0x2df0; same arithmetic as ArrayHelpers.hh:66-99).
 * Returns 0, or 1 when a pivot is <= 0 (the reference prints the matrix and throws). */
static int gi_cholesky(double* G, int n, double* bad_pivot) {
  for (int i = 0; i < n; i++) {
    for (int j = i; j < n; j++) {
      double sum = G[i * n + j];
      for (int k = i - 1; k >= 0; k--) sum -= G[i * n + k] * G[j * n + k];
      if (i == j) {
        if (sum <= 0.0) {
          *bad_pivot = sum; /* the value the reference prints in its exception text */
          return 1;
        }
        G[i * n + i] = sqrt(sum);
      } else {
        G[j * n + i] = sum / G[i * n + i];
      }
    }
    for (int k = i + 1; k < n; k++) G[i * n + k] = G[k * n + i];
  }
  return 0;
}

/* forward_elimination / backward_elimination (ArrayHelpers.hh:174-200). */
static void gi_forward(const double* L, double* y, const double* b, int n) {
  y[0] = b[0] / L[0];
  for (int i = 1; i < n; i++) {
    y[i] = b[i];
    for (int j = 0; j < i; j++) y[i] -= L[i * n + j] * y[j];
    y[i] = y[i] / L[i * n + i];
  }
}
static void gi_backward(const double* U, double* x, const double* y, int n) {
  x[n - 1] = y[n - 1] / U[(n - 1) * n + (n - 1)];
  for (int i = n - 2; i >= 0; i--) {
    x[i] = y[i];
    for (int j = i + 1; j < n; j++) x[i] -= U[i * n + j] * x[j];
    x[i] = x[i] / U[i * n + i];
  }
}

/* compute_d (ArrayHelpers.hh:239-252): d = J^T np. */
static void gi_compute_d(double* d, const double* J, const double* np, int n) {
  for (int i = 0; i < n; i++) {
    double sum = 0.0;
    for (int j = 0; j < n; j++) sum += J[j * n + i] * np[j];
    d[i] = sum;
  }
}
/* update_z (ArrayHelpers.hh:256-267): z = J[:, iq:] d[iq:]. */
static void gi_update_z(double* z, const double* J, const double* d, int iq, int n) {
  for (int i = 0; i < n; i++) {
    z[i] = 0.0;
    for (int j = iq; j < n; j++) z[i] += J[i * n + j] * d[j];
  }
}
/* update_r (ArrayHelpers.hh:271-284): r = R[0:iq,0:iq]^-1 d[0:iq]. */
static void gi_update_r(const double* R, double* r, const double* d, int iq, int n) {
  for (int i = iq - 1; i >= 0; i--) {
    double sum = 0.0;
    for (int j = i + 1; j < iq; j++) sum += R[i * n + j] * r[j];
    r[i] = (d[i] - sum) / R[i * n + i];
  }
}

/* add_constraint (QuadProg++.o+0x21fd). */
static int gi_add_constraint(double* R, double* J, double* d, int* iq_io, double* R_norm, int n) {
  int iq = *iq_io;
  for (int j = n - 1; j >= iq + 1; j--) {
    double cc = d[j - 1], ss = d[j];
    double h = gi_distance(cc, ss);
    if (fabs(h) < GI_EPS) continue;
    d[j] = 0.0;
    ss = ss / h;
    cc = cc / h;
    if (cc < 0.0) {
      cc = -cc;
      ss = -ss;
      d[j - 1] = -h;
    } else {
      d[j - 1] = h;
    }
    double xny = ss / (1.0 + cc);
    for (int k = 0; k < n; k++) {
      double t1 = J[k * n + j - 1], t2 = J[k * n + j];
      J[k * n + j - 1] = t1 * cc + t2 * ss;
      J[k * n + j] = xny * (t1 + J[k * n + j - 1]) - t2;
    }
  }
  iq++;
  for (int i = 0; i < iq; i++) R[i * n + iq - 1] = d[i];
  *iq_io = iq;
  if (fabs(d[iq - 1]) <= GI_EPS * *R_norm) return 0;
  if (*R_norm < fabs(d[iq - 1])) *R_norm = fabs(d[iq - 1]); /* std::max */
  return 1;
}

/* delete_constraint (QuadProg++.o+0x26a8). */
static void gi_delete_constraint(double* R, double* J, int* A, double* u, int n, int p, int* iq_io,
                                 int l) {
  int iq = *iq_io, qq = 0;
  for (int i = p; i < iq; i++)
    if (A[i] == l) {
      qq = i;
      break;
    }
  for (int i = qq; i < iq - 1; i++) {
    A[i] = A[i + 1];
    u[i] = u[i + 1];
    for (int j = 0; j < n; j++) R[j * n + i] = R[j * n + i + 1];
  }
  A[iq - 1] = A[iq];
  u[iq - 1] = u[iq];
  A[iq] = 0;
  u[iq] = 0.0;
  for (int j = 0; j < iq; j++) R[j * n + iq - 1] = 0.0;
  iq--;
  *iq_io = iq;
  if (iq == 0) return;
  for (int j = qq; j < iq; j++) {
    double cc = R[j * n + j], ss = R[(j + 1) * n + j];
    double h = gi_distance(cc, ss);
    if (fabs(h) < GI_EPS) continue;
    cc = cc / h;
    ss = ss / h;
    R[(j + 1) * n + j] = 0.0;
    if (cc < 0.0) {
      R[j * n + j] = -h;
      cc = -cc;
      ss = -ss;
    } else {
      R[j * n + j] = h;
    }
    double xny = ss / (1.0 + cc);
    for (int k = j + 1; k < iq; k++) {
      double t1 = R[j * n + k], t2 = R[(j + 1) * n + k];
      R[j * n + k] = t1 * cc + t2 * ss;
      R[(j + 1) * n + k] = xny * (t1 + R[j * n + k]) - t2;
    }
    for (int k = 0; k < n; k++) {
      double t1 = J[k * n + j], t2 = J[k * n + j + 1];
      J[k * n + j] = t1 * cc + t2 * ss;
      J[k * n + j + 1] = xny * (J[k * n + j] + t1) - t2;
    }
  }
}

int gi_oracle_solve(int n, int p, int m, double* G, const double* g0, const double* CE,
                    const double* ce0, const double* CI, const double* ci0, double* x, double* f,
                    int* iq_out, int* A_out, int* iters_out, int max_iter) {
  const double inf = INFINITY;
  const int mp = m + p;
  int status = GI_OK, steps = 0;
  /* one block for all scratch; +1 slots because the reference reads A[iq]/u[iq] at iq == m+p */
  size_t nd = (size_t)2 * n * n + 4 * (size_t)(mp + 1) + 4 * (size_t)n;
  double* wd = (double*)calloc(nd ? nd : 1, sizeof(double));
  int* wi = (int*)calloc((size_t)3 * (mp + 1) + (size_t)m + 1, sizeof(int));
  double *R = wd, *J = R + (size_t)n * n, *s = J + (size_t)n * n, *r = s + mp + 1, *u = r + mp + 1,
         *u_old = u + mp + 1, *z = u_old + mp + 1, *d = z + n, *np = d + n, *x_old = np + n;
  int *A = wi, *A_old = A + mp + 1, *iai = A_old + mp + 1, *iaexcl = iai + mp + 1;
  double f_value = 0.0, c1, c2, R_norm, psi, ss, sum, t, t1, t2;
  int iq = 0, ip = 0, l = 0;

  if (n <= 0) goto done;

  c1 = 0.0;
  for (int i = 0; i < n; i++) c1 += G[i * n + i];
  if (gi_cholesky(G, n, &f_value)) {
    status = GI_NOT_PD; /* f reports the non-positive pivot ("Error in cholesky decomposition, sum: ...") */
    goto done;
  }
  R_norm = 1.0;
  c2 = 0.0;
  for (int i = 0; i < n; i++) {
    d[i] = 1.0;
    gi_forward(G, z, d, n);
    for (int j = 0; j < n; j++) J[i * n + j] = z[j];
    c2 += z[i];
    d[i] = 0.0;
  }
  /* x = -G^-1 g0 (cholesky_solve, QuadProg++.o+0x31a2); f = 0.5 g0.x */
  gi_forward(G, z, g0, n); /* z doubles as the intermediate y */
  gi_backward(G, x, z, n);
  for (int i = 0; i < n; i++) x[i] = -x[i];
  f_value = 0.5 * gi_dot(g0, x, n);

  /* equality constraints */
  for (int i = 0; i < p; i++) {
    for (int j = 0; j < n; j++) np[j] = CE[j * p + i];
    gi_compute_d(d, J, np, n);
    gi_update_z(z, J, d, iq, n);
    gi_update_r(R, r, d, iq, n);
    t2 = 0.0;
    if (fabs(gi_dot(z, z, n)) > GI_EPS) t2 = (-gi_dot(np, x, n) - ce0[i]) / gi_dot(z, np, n);
    for (int k = 0; k < n; k++) x[k] += t2 * z[k];
    u[iq] = t2;
    for (int k = 0; k < iq; k++) u[k] -= t2 * r[k];
    f_value += 0.5 * (t2 * t2) * gi_dot(z, np, n);
    A[i] = -i - 1;
    if (!gi_add_constraint(R, J, d, &iq, &R_norm, n)) {
      status = GI_EQ_DEPENDENT;
      goto done;
    }
  }
  for (int i = 0; i < m; i++) iai[i] = i;

l1:
  for (int i = p; i < iq; i++) iai[A[i]] = -1;
  ss = 0.0;
  psi = 0.0;
  ip = 0;
  for (int i = 0; i < m; i++) {
    iaexcl[i] = 1;
    sum = 0.0;
    for (int j = 0; j < n; j++) sum += CI[j * m + i] * x[j];
    sum += ci0[i];
    s[i] = sum;
    psi += (sum < 0.0) ? sum : 0.0; /* std::min(0.0, sum) */
  }
  if (fabs(psi) <= (double)m * GI_EPS * c1 * c2 * 100.0) goto done;
  for (int i = 0; i < iq; i++) {
    u_old[i] = u[i];
    A_old[i] = A[i];
  }
  for (int i = 0; i < n; i++) x_old[i] = x[i];

l2:
  for (int i = 0; i < m; i++) {
    if (s[i] < ss && iai[i] != -1 && iaexcl[i]) {
      ss = s[i];
      ip = i;
    }
  }
  if (ss >= 0.0) goto done;
  for (int i = 0; i < n; i++) np[i] = CI[i * m + ip];
  u[iq] = 0.0;
  A[iq] = ip;

l2a:
  steps++;
  if (max_iter > 0 && steps > max_iter) {
    status = GI_ITER_CAP;
    goto done;
  }
  gi_compute_d(d, J, np, n);
  gi_update_z(z, J, d, iq, n);
  gi_update_r(R, r, d, iq, n);
  l = 0;
  t1 = inf;
  for (int k = p; k < iq; k++) {
    if (r[k] > 0.0) {
      if (u[k] / r[k] < t1) {
        t1 = u[k] / r[k];
        l = A[k];
      }
    }
  }
  if (fabs(gi_dot(z, z, n)) > GI_EPS) {
    t2 = -s[ip] / gi_dot(z, np, n);
    if (t2 < 0) t2 = inf;
  } else {
    t2 = inf;
  }
  t = t1 < t2 ? t1 : t2; /* std::min(t1, t2) */
  if (t >= inf) {
    status = GI_INFEASIBLE;
    f_value = inf;
    goto done;
  }
  if (t2 >= inf) {
    for (int k = 0; k < iq; k++) u[k] -= t * r[k];
    u[iq] += t;
    iai[l] = l;
    gi_delete_constraint(R, J, A, u, n, p, &iq, l);
    goto l2a;
  }
  for (int k = 0; k < n; k++) x[k] += t * z[k];
  f_value += t * gi_dot(z, np, n) * (0.5 * t + u[iq]);
  for (int k = 0; k < iq; k++) u[k] -= t * r[k];
  u[iq] += t;
  if (fabs(t - t2) < GI_EPS) {
    if (!gi_add_constraint(R, J, d, &iq, &R_norm, n)) {
      gi_oracle_degenerate_count++;
      iaexcl[ip] = 0;
      gi_delete_constraint(R, J, A, u, n, p, &iq, ip);
      for (int i = 0; i < m; i++) iai[i] = i;
      for (int i = p; i < iq; i++) {
        A[i] = A_old[i];
        u[i] = u_old[i];
        iai[A[i]] = -1;
      }
      for (int i = 0; i < n; i++) x[i] = x_old[i];
      goto l2;
    } else {
      iai[ip] = -1;
    }
    goto l1;
  }
  iai[l] = l;
  gi_delete_constraint(R, J, A, u, n, p, &iq, l);
  sum = 0.0;
  for (int k = 0; k < n; k++) sum += CI[k * m + ip] * x[k];
  s[ip] = sum + ci0[ip];
  goto l2a;

done:
  *f = f_value;
  if (iq_out) *iq_out = iq;
  if (A_out) {
    for (int i = 0; i < mp; i++) A_out[i] = i < iq ? A[i] : 0x7fffffff;
  }
  if (iters_out) *iters_out = steps;
  free(wd);
  free(wi);
  return status;
}

/* ---------------------------------------------------------------- threaded batch driver */
typedef struct {
  int n, p, m, max_iter;
  long long batch;
  const double *G, *g0, *CE, *ce0, *CI, *ci0;
  double *x, *f;
  int *status, *iq, *A, *iters;
  long long* next; /* shared ticket */
  pthread_mutex_t* mu;
} gi_job;

static void* gi_worker(void* arg) {
  gi_job* jb = (gi_job*)arg;
  const int n = jb->n, p = jb->p, m = jb->m;
  const long long chunk = 256;
  double* Gc = (double*)malloc(sizeof(double) * (size_t)(n * n > 0 ? n * n : 1));
  for (;;) {
    pthread_mutex_lock(jb->mu);
    long long b0 = *jb->next;
    *jb->next = b0 + chunk;
    pthread_mutex_unlock(jb->mu);
    if (b0 >= jb->batch) break;
    long long b1 = b0 + chunk < jb->batch ? b0 + chunk : jb->batch;
    for (long long b = b0; b < b1; b++) {
      memcpy(Gc, jb->G + b * n * n, sizeof(double) * (size_t)n * n);
      jb->status[b] = gi_oracle_solve(
          n, p, m, Gc, jb->g0 + b * n, jb->CE + b * n * p, jb->ce0 + b * p, jb->CI + b * n * m,
          jb->ci0 + b * m, jb->x + b * n, jb->f + b, jb->iq ? jb->iq + b : NULL,
          jb->A ? jb->A + b * (p + m) : NULL, jb->iters ? jb->iters + b : NULL, jb->max_iter);
    }
  }
  free(Gc);
  return NULL;
}

int gi_oracle_solve_batch(int n, int p, int m, long long batch, const double* G, const double* g0,
                          const double* CE, const double* ce0, const double* CI, const double* ci0,
                          double* x, double* f, int* status, int* iq, int* A, int* iters,
                          int max_iter, int threads) {
  long long next = 0;
  pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
  gi_job jb = {n, p, m, max_iter, batch, G, g0, CE, ce0, CI, ci0, x, f, status, iq, A, iters, &next, &mu};
  if (threads <= 1) {
    gi_worker(&jb);
    return 0;
  }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
  for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, gi_worker, &jb);
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  free(th);
  return 0;
}
