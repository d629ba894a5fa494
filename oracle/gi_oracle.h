/* TEST INFRASTRUCTURE ONLY -- see gi_oracle.c. */
#ifndef GI_ORACLE_H
#define GI_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes shared with include/mgqp_b200.h (MGQP_STATUS_*). */
enum {
  GI_OK = 0,
  GI_INFEASIBLE = 1,   /* reference returns +inf                                   */
  GI_NOT_PD = 2,       /* reference throws std::logic_error from the Cholesky      */
  GI_EQ_DEPENDENT = 3, /* reference throws std::runtime_error("... dependent")     */
  GI_ITER_CAP = 4      /* reference would not terminate (SURVEY.md App. A note N3) */
};

/* One QP, QuadProg++ layout (G n*n, CE n*p, CI n*m, row-major, constraint = column).
 * G is overwritten by its Cholesky factor exactly like the reference does.
 * A_out (length p+m), iq_out, iters_out may be NULL.  max_iter <= 0 means "no cap". */
int gi_oracle_solve(int n, int p, int m, double* G, const double* g0, const double* CE,
                    const double* ce0, const double* CI, const double* ci0, double* x, double* f,
                    int* iq_out, int* A_out, int* iters_out, int max_iter);

/* Batch over contiguous SoA buffers, `threads` host threads.  G is NOT modified (copied).
 * A is [batch][p+m]; unused tail entries are written as 0x7fffffff. */
int gi_oracle_solve_batch(int n, int p, int m, long long batch, const double* G, const double* g0,
                          const double* CE, const double* ce0, const double* CI, const double* ci0,
                          double* x, double* f, int* status, int* iq, int* A, int* iters,
                          int max_iter, int threads);

#ifdef __cplusplus
}
#endif
#endif
