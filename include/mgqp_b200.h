/* mgqp_b200.h -- C ABI of the B200-native batched dense QP engine for the mgqp hot path.
 *
 * The reference (erreur404/Motion-generation-using-quadratic-programs) solves every QP of its
 * control cycle through ONE call (src/mgqp.cpp:708 and :725):
 *
 *     double solve_quadprog(Matrix<double>& G, Vector<double>& g0,
 *                           const Matrix<double>& CE, const Vector<double>& ce0,
 *                           const Matrix<double>& CI, const Vector<double>& ci0,
 *                           Vector<double>& x);          // include/QuadProgpp/QuadProg++.hh:69-72
 *
 * implemented by the prebuilt lib/QuadProgpp/libquadprog.a (Goldfarb-Idnani dual active set,
 * fp64).  This library replaces that path:
 *
 *   - include/QuadProgpp/QuadProg++.hh + libmgqp_b200.so export the SAME C++ symbol
 *     (_Z14solve_quadprogRN7ArrayHH6MatrixIdEERNS_6VectorIdEERKS1_RKS4_S7_S9_S5_), so
 *     src/mgqp.cpp links unchanged;
 *   - the entry points below are the batched form the reference's FFI for this path would bind.
 *
 * Problem (QuadProg++.hh:8-13):   min 1/2 x'Gx + g0'x   s.t.  CE'x + ce0 = 0,  CI'x + ci0 >= 0
 * Orientation (QuadProg++.hh:15-25): G n*n, CE n*p, CI n*m, row-major, ONE CONSTRAINT PER COLUMN.
 *
 * Everything is plain pointers and sizes; no C++ or torch types cross this boundary.
 */
#ifndef MGQP_B200_H
#define MGQP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-QP verdicts.  0/1 are the two outcomes the reference RETURNS (cost, or +inf:
 * QuadProg++.hh:27-29, tested by the caller at src/mgqp.cpp:717); 2/3 are the two it THROWS
 * (std::logic_error from cholesky_decomposition, std::runtime_error "Constraints are linearly
 * dependent"); 4 does not exist in the reference, which has no iteration limit and can spin
 * forever after a degenerate add_constraint (note N3). */
enum {
  MGQP_STATUS_OK = 0,
  MGQP_STATUS_INFEASIBLE = 1,   /* f = +inf, x = last iterate (as the reference leaves it)    */
  MGQP_STATUS_NOT_PD = 2,       /* f = the non-positive Cholesky pivot (for the error message) */
  MGQP_STATUS_EQ_DEPENDENT = 3, /* f, x = state when the dependent equality was met            */
  MGQP_STATUS_ITER_CAP = 4,
  MGQP_STATUS_RETRIED = 16 /* bit, with MGQP_FLAG_RETRY_WITHOUT_CI only: the low bits are the retry's verdict */
};

/* Return codes of the entry points themselves (argument / CUDA errors; never a per-QP verdict). */
enum {
  MGQP_OK = 0,
  MGQP_ERR_INVALID_ARG = -1,
  MGQP_ERR_UNSUPPORTED_SHAPE = -2,
  MGQP_ERR_CUDA = -3,
  MGQP_ERR_NO_DEVICE = -4
};

/* flags */
enum {
  MGQP_FLAG_DEFAULT = 0,
  /* Allow fused multiply-add and tree reductions.  Faster, but x/f are then only within ~1e-12 of
   * the reference instead of bit-identical, and a decision that sits exactly on a tie can differ.
   * Off by default: the default path reproduces the reference's arithmetic operation by operation. */
  MGQP_FLAG_FAST_MATH = 1,
  /* Stage inputs with plain coalesced loads instead of TMA bulk copies (debug / odd alignments). */
  MGQP_FLAG_NO_BULK_COPY = 2,
  /* Static round-robin tile assignment instead of the persistent atomic work queue. */
  MGQP_FLAG_STATIC_SCHEDULE = 4,
  /* Always use the lane-group kernel (8/16/32 lanes per QP), even where the one-QP-per-lane
   * kernel (n <= 8, m <= 64) would be picked. */
  MGQP_FLAG_GROUP_KERNEL = 8,
  /* Always use the CTA-per-QP kernel (the one that serves n > 32). */
  MGQP_FLAG_CTA_KERNEL = 16,
  /* Host-buffer entry points: ship CI to the device as it is (dense) instead of in the
   * column-compressed transport form described at mgqp_batch::CI_val.  Device-resident entry
   * point: do not classify a dense CI on the device either (n in 9..32: the lane-group kernel
   * otherwise keeps a block made of bound columns compressed in shared memory; same answers). */
  MGQP_FLAG_NO_HOST_PACK = 32,
  /* The reference adapter's policy for an infeasible answer (solveNextStep, src/mgqp.cpp:717-735,
   * TRY_TO_CONVERGE): every QP whose cost comes back NaN or +inf is solved again WITHOUT its
   * inequalities, on the G the first call left behind -- its Cholesky factor, mirrored
   * (QuadProg++.hh:42-45; the identity for the reference's own G = I) -- and if that answer is
   * NaN / +inf too, x is zeroed (:732).  Such QPs come back with MGQP_STATUS_RETRIED or-ed into
   * their status (16 = solved by the retry, 17 = lost twice, x = 0; 18 / 19 = the retry ran into
   * one of the reference's exceptions).  The call stays asynchronous: the second launch reads the
   * number of QPs that go round again from device memory. */
  MGQP_FLAG_RETRY_WITHOUT_CI = 64
};

/* One batch of `batch` independent QPs of one shape, field-wise SoA, DEVICE pointers.
 * Field k of QP b starts at  ptr + b * stride  (stride in ELEMENTS; 0 broadcasts one instance to
 * the whole batch; the natural stride -- n*n, n, n*p, p, n*m, m -- enables the TMA staging path).
 * Within a QP the layout is exactly QuadProg++'s row-major Matrix storage (Array.hh:910-919), so
 * B = 1 is a plain copy of Matrix::v[0]. */
typedef struct mgqp_batch {
  int32_t n, p, m;
  int64_t batch;
  const double *G, *g0, *CE, *ce0, *CI, *ci0; /* inputs are never modified                       */
  int64_t sG, sg0, sCE, sce0, sCI, sci0;
  double* x;       /* [batch][n]                                                              */
  double* f;       /* [batch]    cost, +inf if infeasible                                     */
  int32_t* status; /* [batch]    MGQP_STATUS_*                                                */
  int32_t* iq;     /* [batch]    size of the final active set (equalities included); or NULL  */
  int32_t* A;      /* [batch][p+m] active set in the solver's encoding: equality i -> -(i+1),
                      inequality i -> i; first iq[b] entries valid, rest 0x7fffffff; or NULL   */
  int32_t* iters;  /* [batch]    number of step-direction computations (l2a passes); or NULL  */
  int32_t max_iter; /* <= 0: default 100 + 10*(p+m)                                            */
  int32_t flags;    /* MGQP_FLAG_*                                                              */
  /* Optional COLUMN-COMPRESSED form of CI, for batches in which every inequality has at most one
   * nonzero coefficient -- a bound on one variable, which is every inequality the reference itself
   * builds (limitsMatrix = [-I; +I], src/mgqp.cpp:1080-1128).  When CI_val != NULL, CI / sCI are
   * ignored and column i of QP b is  CI_val[b*sCIv + i] * e_k  with  k = CI_row[b*sCIr + i] < n
   * (an all-zero column: value 0, any row).  m*9 bytes per QP instead of n*m*8.  Served by the
   * one-QP-per-lane and lane-group kernels (n <= 32); larger n returns MGQP_ERR_UNSUPPORTED_SHAPE.
   * The host-buffer entry points build this form themselves, chunk by chunk, on host threads
   * (mgqp_pack_bound_columns) whenever a chunk qualifies, so that the PCIe link carries 1.2 KB
   * instead of 3.2 KB per 7-DoF QP; results are bit-identical either way. */
  const double* CI_val;
  const uint8_t* CI_row;
  int64_t sCIv, sCIr; /* element strides between QPs (>= m); 0 broadcasts */
} mgqp_batch;

/* Solve a device-resident batch on `device`, asynchronously on `stream` (a cudaStream_t passed
 * as void*; NULL = the legacy default stream).  Thread-safe; no global mutable state apart from a
 * per-device ring of work-queue counters.  Replaces QuadProg++.hh:69-72 for batches. */
int solve_quadprog_batched(const mgqp_batch* d, int device, void* stream);

/* Same, but every pointer in `h` is a HOST pointer (pinned or pageable).  The call uploads the
 * inputs in chunks, solves, downloads x/f/status(/iq/A/iters) and returns when the results are in
 * host memory; copies overlap compute through two streams.  This is the "plugin call with host
 * buffers" that bench.py times as e2e.  chunk <= 0 picks a default. */
int solve_quadprog_batched_host(const mgqp_batch* h, int device, int64_t chunk);

/* Host buffers, sharded by QP index over `ndev` devices (devices[0..ndev-1]) with one host thread
 * and its own streams per device, chunks handed out from a shared atomic counter
 * (no collective, the QPs are independent). */
int solve_quadprog_batched_multi(const mgqp_batch* h, const int* devices, int ndev, int64_t chunk);

/* Single QP with host pointers in QuadProg++ layout; what the C++ drop-in solve_quadprog() calls.
 * G_factor (optional, n*n) receives the Cholesky factor the reference leaves in G
 * (QuadProg++.hh:42-45).  Returns MGQP_OK and the verdict through *status. */
int mgqp_solve_one(int32_t n, int32_t p, int32_t m, const double* G, const double* g0,
                   const double* CE, const double* ce0, const double* CI, const double* ci0,
                   double* x, double* f, int32_t* status, double* G_factor, int device);

/* Column-compress the CI blocks of `count` QPs (host memory, QuadProg++ layout n*m, element stride
 * sCI between QPs) into val[count][sv] / row[count][sr] as described at mgqp_batch::CI_val.
 * Returns 1 if every column of every QP has at most one nonzero coefficient (the output is then
 * valid), 0 otherwise.  Pure host code, re-entrant; `threads` <= 0 picks the host's core count. */
int mgqp_pack_bound_columns(int32_t n, int32_t m, int64_t count, const double* CI, int64_t sCI,
                            double* val, int64_t sv, uint8_t* row, int64_t sr, int threads);

/* The same for chunks in which every nonzero coefficient is exactly +1.0 or -1.0 -- every inequality the reference's
 * caller builds (limitsMatrix = [-I; +I], src/mgqp.cpp:1111-1112): ONE code byte per column (bits 0-5 the row, bit 7
 * set for -1.0, 0x40 for an all-zero column), code[count][sc].  Returns 1 if the chunk has that form, 0 otherwise.
 * The host-buffer entry points try this form first (the link then carries 0.9 KB per 7-DoF QP) and expand the codes
 * into (value, row) pairs on the device; a chunk that does not qualify goes through mgqp_pack_bound_columns.
 * mgqp_pack_unit_is_fast: 1 iff this CPU has the vector (AVX-512) version; without it the form is not used. */
int mgqp_pack_unit_columns(int32_t n, int32_t m, int64_t count, const double* CI, int64_t sCI, uint8_t* code, int64_t sc,
                           int threads);
int mgqp_pack_unit_is_fast(void);

/* Host threads the packer uses by default: MGQP_HOST_THREADS, else the cores this process may run on.
 * The host-buffer entry points pack CI only when these threads can outrun what the packing saves on
 * the link (one rank per GPU on a 32-core box leaves 4 threads per GPU: dense transport wins there). */
int mgqp_host_pack_threads(void);

/* 1 iff all `count` blocks of `size` doubles (element stride `stride`) equal the first one bit for bit.
 * The host-buffer entry points use it on G and g0, chunk by chunk: the reference's own caller passes
 * G = I and g0 = 0 for every QP (src/mgqp.cpp:672, :706), and such a field is then uploaded once per
 * chunk (device-side stride 0) instead of once per QP. */
int mgqp_blocks_all_equal(const double* base, int64_t stride, int64_t size, int64_t count, int threads);

/* Bytes the last host-buffer call of the calling thread moved over the host<->device link. */
void mgqp_last_transfer_bytes(int64_t* h2d, int64_t* d2h);

/* Chunks each device took in the calling thread's last solve_quadprog_batched_multi (the shared queue balances
 * uneven work, e.g. the infeasible instances of a scenario sweep); returns the number of devices of that call. */
int mgqp_last_multi_chunks(int64_t* counts, int max_devices);

/* Self-test of the lane kernel's split division and square root (csrc/warp_prims.cuh: d_rcp_refined / d_div_rcp /
 * d_sqrt_split, the inline paths of div.rn.f64 and sqrt.rn.f64 computed branch-free): `count` operand pairs drawn
 * from `seed` (raw bit patterns over the whole exponent range, operands of similar magnitude, pairs sharing a
 * divisor) on the current device.  accepted = operations the split paths answer themselves; mismatches = accepted
 * operations whose result differs in any bit from the compiler's own division / square root (must be 0).
 * Returns MGQP_OK or the CUDA failure. */
int mgqp_debug_div_check(int64_t count, uint64_t seed, int64_t* accepted, int64_t* mismatches);

/* Introspection. */
const char* mgqp_status_string(int status);
const char* mgqp_last_error(void);        /* thread-local text of the last non-OK return        */
int mgqp_device_count(void);
/* Kernel launches issued by this library in this process (all threads); for bench bookkeeping. */
int64_t mgqp_kernel_launch_count(void);
/* Which kernel family a shape maps to: 1 = one QP per lane, 8/16/32 = lanes per QP of the warp kernel, 0 = CTA-per-QP
 * kernel, <0 = unsupported.  smem_bytes_per_cta / ctas_per_sm (optional) receive the launch plan. */
int mgqp_plan(int32_t n, int32_t p, int32_t m, int32_t* smem_bytes_per_cta, int32_t* warps_per_cta);
/* Compulsory HBM bytes per QP for a shape (every input once, x/f/status once); SURVEY.md 8(d). */
int64_t mgqp_algorithmic_bytes(int32_t n, int32_t p, int32_t m);

/* ---- The solver's two callers, batched, for the 7-DoF wiring the reference ships
 * (ops/mgqp.ops:233-240, :256, :289-292): per scenario the QP assembly of updateHook (src/mgqp.cpp:886-1143) and the
 * priority cascade of solveNextHierarchy (:751-869) -- level 0 = translational task of link 7 + dynamics [M -I] under
 * the box limits [-I; +I] with log-barrier shaped acceleration limits, level 2 = joint-1 position task in level 0's
 * null space -- each level solved by the engine above with solveNextStep's infeasible-retry policy (:717-735).
 * Inputs are the ~0.5 KB of robot state the reference reads from its ports (float32, SoA, [batch][...] row-major),
 * outputs the 14-vector [qdd; tau] solveNextHierarchy returns and the torque command tau + h (:1152). */
typedef struct mgqp_scenarios {
  int64_t batch;
  const float *q, *qd;           /* [batch][7]   joint angles / velocities (in_robotstatus)          */
  const float* M;                /* [batch][7][7] joint-space inertia (in_inertia)                    */
  const float* h;                /* [batch][7]   Coriolis + gravity (in_h)                            */
  const float *J, *Jd;           /* [batch][3][7] translation Jacobian of link 7 and its derivative   */
  const float *pos_des, *pos_cur, *vel_des, *vel_cur, *acc_des, *acc_cur; /* [batch][3] task-space targets / state */
  const float* q1_des;           /* [batch]      desired position of the joint-space task's joint     */
  float kp_task, kd_task, kp_joint, kd_joint;          /* gains (src/mgqp.cpp:104-110)                 */
  float acc_p[7], acc_n[7], tau_p[7], tau_n[7], q_sup[7], q_inf[7]; /* limits (ops/mgqp.ops:187-189)  */
  int32_t joint_task;            /* 0-based joint of the level-2 task (shipped: 0)                    */
  float* tracking;               /* [batch][14]  [qdd; tau] (zeros where level 0 failed)              */
  float* torques;                /* [batch][7]   tau + h                                              */
  int32_t *status0, *status2;    /* [batch] verdicts of the two levels as MGQP_FLAG_RETRY_WITHOUT_CI reports them
                                    (status2 = -1 where level 0 failed); optional                       */
  /* optional taps for tests: the level-0 QP as assembled (fp64, solver layout) and the projector */
  double *tap_CE0, *tap_ce0, *tap_ci0; /* [batch][14][10], [batch][10], [batch][28] */
  float* tap_Z;                        /* [batch][14][14] */
} mgqp_scenarios;

/* Device pointers; asynchronous on `stream`. */
int mgqp_cascade_batched(const mgqp_scenarios* s, int device, void* stream);
/* Host pointers: state uploaded / answers downloaded chunk by chunk over two streams. */
int mgqp_cascade_batched_host(const mgqp_scenarios* s, int device, int64_t chunk);

/* Bench hook (BASELINE.json configs[0], the single-QP-per-cycle path): wall-clock microseconds of `reps`
 * consecutive calls of the C++ drop-in solve_quadprog() (QuadProg++.hh:69-72) on one QP in QuadProg++ layout.
 * us_out[reps]; x_out[n] / f_out (optional) receive the last call's answer.  Returns 0, -1 if a call threw. */
int mgqp_dropin_latency(int32_t n, int32_t p, int32_t m, const double* G, const double* g0, const double* CE,
                        const double* ce0, const double* CI, const double* ci0, int reps, double* us_out,
                        double* x_out, double* f_out);

#ifdef __cplusplus
}
#endif
#endif /* MGQP_B200_H */
