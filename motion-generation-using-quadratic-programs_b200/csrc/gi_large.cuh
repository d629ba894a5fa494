// Goldfarb-Idnani dual active-set solver, ONE QP PER CTA (n > 32, or m too large for the warp
// kernels): the horizon-stacked shapes (BASELINE.json configs[3]: n = 350, m = 2100).
//
// Replaces the body of the reference's solve_quadprog (lib/QuadProgpp/libquadprog.a: QuadProg++.o;
// prototype include/QuadProgpp/QuadProg++.hh:69-72; spec SURVEY.md Appendix A).  At this size the
// matrices no longer fit on chip: the Cholesky factor, J (stored TRANSPOSED, so that update_z and the
// Givens sweeps -- thread k works on row k of J -- are coalesced) and R (n*n doubles each) live in a
// per-CTA global-memory workspace that stays L2-resident between uses, the vectors and the per-rotation
// Givens coefficients live in shared memory, and CE/CI are read in place (CI once per
// violated-constraint scan, which is what bounds this kernel: 8*n*m bytes per iteration).
//
// Parallel decomposition (T threads, every reduction order is the reference's):
//   * every dot product of length n is owned by ONE thread and accumulated in ascending index
//     order, one rounding per operation: the threads split the OUTPUT index (column i of the scan,
//     element i of d = J^T np and z = J d), never the summation index;
//   * add_constraint's Givens sweep is split in three: a serial chain that only produces the
//     hypotenuses h_j (`distance` ignores the sign of its arguments, so the chain needs nothing
//     else), one parallel pass that turns them into (c, s, s/(1+c)) with the reference's sign
//     normalisation, and a row-parallel pass in which each thread streams its own rows of J right to
//     left through all rotations -- no barrier between rotations;
//   * the triangular solves whose subtraction order is the reverse of data availability
//     (backward_elimination, update_r) are left to one thread.
// Control flow is uniform across the CTA: thread 0's decisions are broadcast through shared memory.
#pragma once
#include "gi_lane.cuh"  // LaneIO, ln_distance, MGQP_LDG, MGQP_EPS

#ifndef MGQP_LARGE_GB
#define MGQP_LARGE_GB 8  // columns of J fetched ahead in add_constraint's row pass (at 80 registers: 8 -> 16 -2 %, 32 -12 %)
#endif
#ifndef MGQP_LARGE_SPLIT_CHAIN
#define MGQP_LARGE_SPLIT_CHAIN 1
#endif
#ifndef MGQP_EMULATE
#include "gi_large_dmma.cuh"
#endif

#ifdef MGQP_EMULATE
#define MGQP_UNROLL8
#define MGQP_UNROLL16
#define MGQP_UNROLL4
#define MGQP_UNROLL_SETUP
#define MGQP_UNROLL_SOLVE
#else
#define MGQP_UNROLL8 _Pragma("unroll 8")
#define MGQP_UNROLL16 _Pragma("unroll 16")
#define MGQP_UNROLL4 _Pragma("unroll 4")
#ifndef MGQP_SETUP_DEPTH
#define MGQP_SETUP_DEPTH 16  // loads in flight per thread in the set-up loops: 8 -> 16 +1.2 %, 32 the same
#endif
#define MGQP_PRAGMA_(x) _Pragma(#x)
#define MGQP_PRAGMA(x) MGQP_PRAGMA_(x)
#define MGQP_UNROLL_SETUP MGQP_PRAGMA(unroll MGQP_SETUP_DEPTH)
// the serial upper-triangular solves (one thread): the deeper the unroll, the better the shared-memory loads and the
// products hide behind the chain of additions
#ifndef MGQP_SOLVE_DEPTH
#define MGQP_SOLVE_DEPTH 16  // 4 -> 16: +1.8 % on C4
#endif
#define MGQP_UNROLL_SOLVE MGQP_PRAGMA(unroll MGQP_SOLVE_DEPTH)
#endif

namespace mgqp {

// rows of R / L staged per block for the serial solves.  1 keeps the CTA's shared memory at 49 KB for the 350-variable
// shape, which is what lets four CTAs share an SM (gi_large_kernel.cuh).
#ifndef MGQP_ROWBLK
#define MGQP_ROWBLK 1
#endif
#ifndef MGQP_ELL_K
#define MGQP_ELL_K 64
#endif
#ifndef MGQP_LARGE_PARHYP
#define MGQP_LARGE_PARHYP 0
#endif

struct LargeView {
  int n, p, m, T;
  // shared memory, in doubles from the (16-byte aligned) base
  int ox, oz, od, onp, orr, ou, oxold, ouold, oh, occ, oss, oxny, ored, obc, orow;
  int onzv;               // doubles: values of the pending column's nonzeros (<= MGQP_ELL_K)
  int oredi, oA, oAold, onzr;   // ints, byte offsets (onzr: rows of the pending column's nonzeros)
  int oact, oexcl, oskip; // bytes, byte offsets
  int smem_bytes;
  int ell_k;              // capacity of the ELL form of CI (nonzeros kept per column)
  long long ws_doubles;   // global workspace per CTA: Gc | J | R | s | ELL values | ELL rows | nnz per column
};

MGQP_HOST inline LargeView make_large_view(int n, int p, int m, int T) {
  LargeView v;
  v.n = n; v.p = p; v.m = m; v.T = T;
  int o = 0;
  v.ox = o; o += n;
  v.oz = o; o += n;
  v.od = o; o += n;
  v.onp = o; o += n;
  v.orr = o; o += n + 1;
  v.ou = o; o += n + 1;
  v.oxold = o; o += n;
  v.ouold = o; o += n + 1;
  v.oh = o; o += n + 1;
  v.occ = o; o += n + 1;
  v.oss = o; o += n + 1;
  v.oxny = o; o += n + 1;
  v.ored = o; o += T;
  v.obc = o; o += 16;
  v.orow = o; o += 2 * MGQP_ROWBLK * n;  // double-buffered block of R / L rows for the serial solves
  v.onzv = o; o += MGQP_ELL_K;
  int b = o * 8;
  v.oredi = b; b += 4 * T;
  v.oA = b; b += 4 * (n + 2);
  v.oAold = b; b += 4 * (n + 2);
  v.onzr = b; b += 4 * MGQP_ELL_K;
  v.oact = b; b += m;
  v.oexcl = b; b += m;
  v.oskip = b; b += n + 1;
  v.smem_bytes = (b + 15) & ~15;
  // + the inequality block in ELL form: up to MGQP_ELL_K (value, row) pairs per column, [k][column] so that the
  //   threads of the scan (one column each) read consecutive addresses; nnz per column behind it
  v.ell_k = n < MGQP_ELL_K ? n : MGQP_ELL_K;
  v.ws_doubles = 3ll * n * n + m + (long long)v.ell_k * m + ((long long)v.ell_k * m + 3) / 4 + (m + 3) / 4 + 2;
  return v;
}

struct LargeOut {
  double* x; double* f; int* status; int* iq; int* A; int* iters;
  // Admission gate of the set-up phase (device build only; null = none).  The factorisation streams its n*n
  // matrices n/3 times; with every resident CTA in that phase at once (hundreds of MB) nothing survives in L2 and
  // each pass comes from DRAM.  At most gate_limit CTAs factorise at a time, so that their matrices stay L2-resident;
  // the others wait here (the SM's second CTA keeps iterating meanwhile).
  unsigned* gate = nullptr;
  int gate_limit = 0;
};

template <bool FMA>
MGQP_DEV void gi_large_solve(const LargeView& v, unsigned char* smem, double* ws, const LaneIO& io, int max_iter,
                             const LargeOut& out) {
  const int n = v.n, p = v.p, m = v.m;
  const int tid = blk_tid(), T = blk_dim();
  double* sm = reinterpret_cast<double*>(smem);
  double* sx = sm + v.ox;
  double* sz = sm + v.oz;
  double* sd = sm + v.od;
  double* snp = sm + v.onp;
  double* sr = sm + v.orr;
  double* su = sm + v.ou;
  double* sxold = sm + v.oxold;
  double* suold = sm + v.ouold;
  double* sh = sm + v.oh;      // hypotenuse per rotation; also scratch
  double* scc = sm + v.occ;
  double* sss = sm + v.oss;
  double* sxny = sm + v.oxny;
  double* sred = sm + v.ored;
  double* sbc = sm + v.obc;    // broadcast slots
  double* srow = sm + v.orow;  // 2 x MGQP_ROWBLK rows
  double* snzv = sm + v.onzv;  // the pending inequality column in sparse form (values, rows): see direction()
  int* snzr = reinterpret_cast<int*>(smem + v.onzr);
  int* sredi = reinterpret_cast<int*>(smem + v.oredi);
  int* A = reinterpret_cast<int*>(smem + v.oA);
  int* Aold = reinterpret_cast<int*>(smem + v.oAold);
  unsigned char* act = smem + v.oact;    // iai[i] == -1
  unsigned char* excl = smem + v.oexcl;  // iaexcl[i]
  unsigned char* skip = smem + v.oskip;  // rotation j skipped (|h| < eps)
  double* Gc = ws;
  double* J = Gc + (size_t)n * n;
  double* R = J + (size_t)n * n;
  double* sg = R + (size_t)n * n;  // s[m], only read back on the rare exact-psi path
  // CI in ELL form (built by the one compulsory pass over the dense block): the scan of an MPC-style constraint
  // matrix -- bounds on accelerations, velocities, positions: identity and integrator blocks, a few percent
  // nonzeros -- then reads the nonzeros instead of 8*n*m bytes per iteration
  const int EK = v.ell_k;
  double* ell_val = sg + m;
  unsigned short* ell_row = reinterpret_cast<unsigned short*>(ell_val + (size_t)EK * m);
  unsigned short* ell_nnz = ell_row + (((size_t)EK * m + 3) & ~(size_t)3);
  constexpr unsigned short ELL_DENSE = 0xffff;  // more than EK nonzeros: this column is scanned densely
  const double INF = d_inf();
  constexpr int NONE = 0x7fffffff;

  int Tp2 = 1;  // T rounded up to a power of two (the block size need not be one: 352 threads give the 350-variable
  while (Tp2 < T) Tp2 <<= 1;  // MPC shape one row per thread)
  // lexicographic (value, index) minimum over the block; every thread gets the result
  auto block_argmin = [&](double val, int idx, double& oval, int& oidx) {
    sred[tid] = val;
    sredi[tid] = idx;
    blk_sync();
    MGQP_UNROLL1
    for (int s = Tp2 >> 1; s > 0; s >>= 1) {
      if (tid < s && tid + s < T) {
        const double a = sred[tid + s];
        const int ai = sredi[tid + s];
        if (a < sred[tid] || (a == sred[tid] && ai < sredi[tid])) {
          sred[tid] = a;
          sredi[tid] = ai;
        }
      }
      blk_sync();
    }
    oval = sred[0];
    oidx = sredi[0];
    blk_sync();
  };
  auto block_sum = [&](double val) -> double {
    sred[tid] = val;
    blk_sync();
    MGQP_UNROLL1
    for (int s = Tp2 >> 1; s > 0; s >>= 1) {
      if (tid < s && tid + s < T) sred[tid] = d_add(sred[tid], sred[tid + s]);
      blk_sync();
    }
    const double r = sred[0];
    blk_sync();
    return r;
  };
  // ascending-order dot product of two shared vectors, done by thread 0 and broadcast
  auto dot0 = [&](const double* a, const double* b, int slot) {
    if (tid == 0) {
      double s = 0.0;
      MGQP_UNROLL1
      for (int i = 0; i < n; i++) s = mad<FMA>(a[i], b[i], s);
      sbc[slot] = s;
    }
  };

  // out[i] = (rhs[i] - sum_{j=i+1}^{k-1} U[i][j]*out[j]) / U[i][i] for i = k-1 .. 0, the sum in
  // ascending j (update_r / backward_elimination).  The recurrence is serial by construction (row i
  // needs out[i+1] before its first addition), so one thread runs it -- but from shared memory:
  // the other threads stream the rows in, MGQP_ROWBLK at a time, one block ahead of the consumer.
  // running == true: backward_elimination's form, a = rhs[i]; a -= U[i][j]*out[j]; out[i] = a / U[i][i].
  auto upper_solve = [&](const double* U, int k, const double* rhs, double* outv, bool running) {
    const int nb = (k + MGQP_ROWBLK - 1) / MGQP_ROWBLK;
    auto fetch = [&](int b, int first, int step) {
      const int hi = k - b * MGQP_ROWBLK;
      const int lo = hi - MGQP_ROWBLK > 0 ? hi - MGQP_ROWBLK : 0;
      double* buf = srow + (size_t)(b & 1) * MGQP_ROWBLK * n;
      MGQP_UNROLL1
      for (int r = lo; r < hi; r++) {
        const double* Ur = U + (size_t)r * n;
        double* br = buf + (size_t)(r - lo) * n;
        MGQP_UNROLL4
        for (int j = r + first; j < k; j += step) br[j] = Ur[j];
      }
    };
    if (nb > 0) fetch(0, tid, T);
    blk_sync();
    MGQP_UNROLL1
    for (int b = 0; b < nb; b++) {
      if (tid == 0) {
        const int hi = k - b * MGQP_ROWBLK;
        const int lo = hi - MGQP_ROWBLK > 0 ? hi - MGQP_ROWBLK : 0;
        const double* buf = srow + (size_t)(b & 1) * MGQP_ROWBLK * n;
        MGQP_UNROLL1
        for (int i = hi - 1; i >= lo; i--) {
          const double* Ui = buf + (size_t)(i - lo) * n;
          if (running) {
            double a = rhs[i];
            MGQP_UNROLL_SOLVE
            for (int j = i + 1; j < k; j++) a = msub<FMA>(a, Ui[j], outv[j]);
            outv[i] = d_div(a, Ui[i]);
          } else {
            double s = 0.0;
            MGQP_UNROLL_SOLVE
            for (int j = i + 1; j < k; j++) s = mad<FMA>(Ui[j], outv[j], s);
            outv[i] = d_div(d_sub(rhs[i], s), Ui[i]);
          }
        }
      } else if (b + 1 < nb) {
        fetch(b + 1, tid - 1, T - 1);
      }
      blk_sync();
    }
  };

  int status = 0, iq = 0, steps = 0;
  double f = 0.0;

  // ---------------------------------------------------------------- set-up
#ifndef MGQP_EMULATE
  if (out.gate) {
    if (tid == 0) {
      while (atomicAdd(out.gate, 1u) >= (unsigned)out.gate_limit) {
        atomicSub(out.gate, 1u);
        __nanosleep(4000);
      }
    }
    blk_sync();
  }
#endif
  bool fail = false;
  double c1 = 0.0, c2 = 0.0, R_norm = 1.0;
  bool dense_done = false;
#ifndef MGQP_EMULATE
  // scratch of the tensor-core set-up: 8 (n + 16) + 64 (warps + 2) doubles -- the row buffer if it is large enough,
  // else the block of n-vectors in front of it (none of them is live before the set-up has finished)
  const int dmma_need = 8 * (n + 16) + 64 * ((T >> 5) + 2);
  double* dmma_scratch = (2 * MGQP_ROWBLK * n >= dmma_need) ? srow : ((v.ored - v.ox >= dmma_need) ? sm + v.ox : nullptr);
  if (FMA && n >= 128 && dmma_scratch != nullptr) {
    // MGQP_FLAG_FAST_MATH: the dense set-up on the fp64 tensor cores (gi_large_dmma.cuh)
    dense_done = true;
    MGQP_UNROLL1
    for (int e = tid; e < n * n; e += T) R[e] = 0.0;
    if (tid == 0) {
      double c = 0.0;
      MGQP_UNROLL1
      for (int i = 0; i < n; i++) c = d_add(c, MGQP_LDG(io.G + (size_t)i * n + i));
      sbc[0] = c;
    }
    blk_sync();
    c1 = sbc[0];
    blk_sync();
    double bad = 0.0;
    if (tid == 0) sbc[2] = 0.0;
    const bool okf = large_setup_dmma(n, io.G, Gc, J, dmma_scratch, &sbc[2]);
    blk_sync();
    if (!okf) {
      fail = true;
      status = 2;
      f = sbc[2];
    } else {
      // y = X g0 (row i of X is contiguous), x0 = -X' y (column j: coalesced), f = 0.5 g0.x0, c2 = trace(J) = trace(X)
      MGQP_UNROLL1
      for (int i = tid; i < n; i += T) {
        double a = 0.0;
        const double* Xi = J + (size_t)i * n;
        MGQP_UNROLL8
        for (int j = 0; j <= i; j++) a = d_fma(Xi[j], MGQP_LDG(io.g0 + j), a);
        sz[i] = a;
      }
      blk_sync();
      MGQP_UNROLL1
      for (int j = tid; j < n; j += T) {
        double a = 0.0;
        MGQP_UNROLL8
        for (int i = j; i < n; i++) a = d_fma(J[(size_t)i * n + j], sz[i], a);
        sx[j] = -a;
      }
      blk_sync();
      if (tid == 0) {
        double t = 0.0, fs = 0.0;
        MGQP_UNROLL1
        for (int i = 0; i < n; i++) {
          t = d_add(t, J[(size_t)i * n + i]);
          fs = d_fma(MGQP_LDG(io.g0 + i), sx[i], fs);
        }
        sbc[0] = t;
        sbc[1] = d_mul(0.5, fs);
      }
      blk_sync();
      c2 = sbc[0];
      f = sbc[1];
      blk_sync();
    }
    (void)bad;
  }
#endif
  bool my_nonfinite = false;  // this thread wrote a factor entry that is inf or NaN
  if (!dense_done) {
  MGQP_UNROLL1
  for (int e = tid; e < n * n; e += T) Gc[e] = MGQP_LDG(io.G + e);
  MGQP_UNROLL1
  for (int e = tid; e < n * n; e += T) R[e] = 0.0;
  blk_sync();
  if (tid == 0) {
    double c = 0.0;
    MGQP_UNROLL1
    for (int i = 0; i < n; i++) c = d_add(c, Gc[(size_t)i * n + i]);
    sbc[0] = c;
  }
  blk_sync();
  c1 = sbc[0];
  blk_sync();

  // Cholesky, row-oriented as the reference; row i is staged in sz for the broadcast operand
  MGQP_UNROLL1
  for (int i = 0; i < n && !fail; i++) {
    MGQP_UNROLL1
    for (int k = tid; k < i; k += T) sz[k] = Gc[(size_t)i * n + k];
    blk_sync();
    MGQP_UNROLL1
    for (int j = i + tid; j < n; j += T) {
      double sum = Gc[(size_t)i * n + j];
      MGQP_UNROLL_SETUP
      // L[j][k] is read as its mirror image Gc[k][j] (written when row k was finished): same bits,
      // and consecutive threads (j) touch consecutive addresses
      for (int k = i - 1; k >= 0; k--) sum = msub<FMA>(sum, sz[k], Gc[(size_t)k * n + j]);
      sd[j] = sum;
    }
    blk_sync();
    const double piv = sd[i];
    if (piv <= 0.0) {
      fail = true;
      status = 2;
      f = piv;
    } else {
      const double dii = d_sqrt(piv);
      MGQP_UNROLL1
      for (int j = i + tid; j < n; j += T) {
        if (j == i) {
          Gc[(size_t)i * n + i] = dii;
          my_nonfinite = my_nonfinite || !(dii < INF);
        } else {
          const double val = d_div(sd[j], dii);
          Gc[(size_t)j * n + i] = val;
          Gc[(size_t)i * n + j] = val;
          my_nonfinite = my_nonfinite || !(d_abs(val) < INF);
        }
      }
    }
    blk_sync();
  }

  }  // !dense_done: Cholesky
  if (!fail && !dense_done) {
    // J = L^-T: thread i owns row i = forward_elimination(L, e_i).  J is upper triangular: with a
    // finite factor the entries left of the diagonal are exactly +0.0 and the products they enter
    // are +-0.0, which never change the accumulator (x - (+-0) = x, and +0.0 - (+-0.0) = +0.0), so the
    // row can start at its diagonal -- a third of the multiply-adds, bit for bit the same result.  A
    // factor with inf/NaN in it (a poisoned G) takes the full loops so that the poison spreads
    // exactly as in the reference.
    if (tid == 0) sredi[0] = 0;
    blk_sync();
    if (my_nonfinite) sredi[0] = 1;
    blk_sync();
    const bool finite_L = sredi[0] == 0;
    blk_sync();
    MGQP_UNROLL1
    for (int ib = 0; ib < n; ib += T) {
      const int i = ib + tid;
      const bool have = i < n;
      double* Ji = J + i;  // row i of J = column i of the transposed store: element r at Ji[r * n]
      // Every lane of a warp runs the same (r, j) ranges, starting at the warp's first row iw: lane i's
      // loads J[j][i] then sit next to its neighbours' (one 256-byte segment per warp), whereas a start
      // at each lane's own diagonal put every lane on a different row j (stride n + 1: one 32-byte
      // sector per lane).  The extra terms j in [iw, i) multiply stored +0.0 entries: no effect.
      const int iw = finite_L ? (ib + (tid & ~31)) : 0;
      if (have && finite_L) {
        MGQP_UNROLL1
        for (int r = 0; r < i; r++) Ji[(size_t)r * n] = 0.0;
      }
      MGQP_UNROLL1
      for (int r = iw; r < n; r++) {
        if (have && (r >= i || !finite_L)) {
          double acc = (r == i) ? 1.0 : 0.0;
          const double* Lr = Gc + (size_t)r * n;
          MGQP_UNROLL_SETUP
          for (int j = iw; j < r; j++) acc = msub<FMA>(acc, Lr[j], Ji[(size_t)j * n]);
          Ji[(size_t)r * n] = d_div(acc, Lr[r]);
        }
      }
    }
    blk_sync();
    // forward elimination for x0 across threads (subtraction order j = 0,1,.. = availability order)
    MGQP_UNROLL1
    for (int i = tid; i < n; i += T) sd[i] = MGQP_LDG(io.g0 + i);
    blk_sync();
    MGQP_UNROLL1
    for (int j = 0; j < n; j++) {
      if (tid == 0) sz[j] = d_div(sd[j], Gc[(size_t)j * n + j]);
      blk_sync();
      const double yj = sz[j];
      MGQP_UNROLL1
      for (int i = j + 1 + tid; i < n; i += T) sd[i] = msub<FMA>(sd[i], Gc[(size_t)i * n + j], yj);
      blk_sync();
    }
    // backward elimination (serial: its ascending-j subtraction is the reverse of availability), c2, f
    upper_solve(Gc, n, sz, sx, true);
    if (tid == 0) {
      double t = 0.0, fs = 0.0;
      MGQP_UNROLL1
      for (int i = 0; i < n; i++) {
        t = d_add(t, J[(size_t)i * n + i]);
        sx[i] = -sx[i];
        fs = mad<FMA>(MGQP_LDG(io.g0 + i), sx[i], fs);
      }
      sbc[0] = t;
      sbc[1] = d_mul(0.5, fs);
    }
    blk_sync();
    c2 = sbc[0];
    f = sbc[1];
    blk_sync();
  }

  // Sparse form of d = J^T np.  d_i = sum_j J[j][i] np[j] is a sum that starts at +0.0; a term with np[j] == 0 is
  // +-0.0 as long as J[j][i] is finite and can then not change a bit of it (a partial sum is never -0.0).  J is
  // stored transposed, thread i walks ITS OWN row of that array: 32 different cache lines per load instruction,
  // the one uncoalesced access pattern of this kernel -- with the few nonzeros of a bound / integrator column
  // (ELL lists, ascending rows = the reference's order) it shrinks from n to nnz loads per thread.
  // j_finite: no entry of J is inf / NaN (tracked wherever J is written; the tensor-core mode does not track).
  int np_nnz = -1;        // >= 0: snzv / snzr hold the pending column's nonzeros
  bool j_bad_mine = false;
  bool j_finite = FMA || !fail;
  if (!FMA && !fail && !dense_done) {
    // J = L^-T as just built: finite iff every entry this thread owns is (rows i = tid, tid + T, ...)
    bool bad = false;
    MGQP_UNROLL1
    for (int i = tid; i < n; i += T) {
      MGQP_UNROLL4
      for (int r = 0; r < n; r++) bad = bad || !(d_abs(J[(size_t)r * n + i]) < INF);
    }
    if (tid == 0) sredi[0] = 0;
    blk_sync();
    if (bad) sredi[0] = 1;
    blk_sync();
    j_finite = sredi[0] == 0;
    blk_sync();
  }
  // d = J^T np ; z = J[:, iq:] d[iq:] ; r = R[0:iq,0:iq]^-1 d[0:iq]
  auto direction = [&]() {
#ifdef MGQP_LARGE_NO_SPARSE_D
    if (false) {
#else
    if (np_nnz >= 0 && j_finite) {
#endif
      MGQP_UNROLL1
      for (int i = tid; i < n; i += T) {
        double s = 0.0;
        const double* Ji = J + (size_t)i * n;
        MGQP_UNROLL8
        for (int k = 0; k < np_nnz; k++) s = mad<FMA>(Ji[snzr[k]], snzv[k], s);
        sd[i] = s;
      }
    } else {
    MGQP_UNROLL1
    for (int i = tid; i < n; i += T) {
      double s = 0.0;
      MGQP_UNROLL16
      for (int j = 0; j < n; j++) s = mad<FMA>(J[(size_t)i * n + j], snp[j], s);  // J[j][i]
      sd[i] = s;
    }
    }
    blk_sync();
    MGQP_UNROLL1
    for (int i = tid; i < n; i += T) {
      double s = 0.0;
      const double* Ji = J + i;
      MGQP_UNROLL16
      for (int j = iq; j < n; j++) s = mad<FMA>(Ji[(size_t)j * n], sd[j], s);  // J[i][j]
      sz[i] = s;
    }
    blk_sync();
    upper_solve(R, iq, sd, sr, false);
  };

  // add_constraint (Appendix A).  Returns the reference's bool; iq is incremented.
  auto add_constraint = [&]() -> bool {
    // (1) hypotenuses only.  sh[j] = h of rotation j, skip[j] = |h| < eps.
    // Optional (-DMGQP_LARGE_PARHYP=1, tensor-core / FMA mode only): h_j = sqrt(sum_{k >= j-1} d[k]^2) as a running
    // sum of squares and parallel square roots instead of one thread's chain of n dependent divisions and square
    // roots.  +14 % on C4 (4.66e3 -> 5.34e3 QP/s), but NOT enabled: measured on 2368 horizon-stacked QPs the iterates
    // then differ from the exact mode's by up to 2e-2 relative (1 of 592 also ends in a different active set) --
    // outside the north-star gate, while the chained `distance` form stays within 1.4e-14.
    bool chain = true;
#if MGQP_LARGE_PARHYP
    if (FMA) {
      bool wide = false;
      MGQP_UNROLL1
      for (int j = tid; j < n; j += T) {
        const double a = d_abs(sd[j]);
        wide = wide || !(a < 1e140) || (a != 0.0 && a < 1e-140);
      }
      if (tid == 0) sredi[2] = 0;
      blk_sync();
      if (wide) sredi[2] = 1;
      blk_sync();
      chain = sredi[2] != 0;
      if (!chain) {
        // S_j = d[j-1]^2 + S_{j+1} as ONE chain of fused multiply-adds (a few thousand cycles, against the chain of
        // divisions and square roots it replaces), so that consecutive hypotenuses stay consistent to one rounding:
        // c^2 + s^2 = (d[j-1]^2 + S_{j+1}) / S_j.  (Independent suffix sums per j drift by ~1e-14 per rotation, and
        // tens of thousands of rotations then cost J its orthogonality: measured 1e-9 in x.)
        if (tid == 0) {
          double acc = (n > 0) ? d_mul(sd[n - 1], sd[n - 1]) : 0.0;
          MGQP_UNROLL4
          for (int j = n - 1; j >= iq + 1; j--) {
            acc = d_fma(sd[j - 1], sd[j - 1], acc);
            sh[j] = acc;
          }
        }
        blk_sync();
        MGQP_UNROLL1
        for (int j = iq + 1 + tid; j <= n - 1; j += T) {
          const double h = d_sqrt(sh[j]);
          skip[j] = (d_abs(h) < MGQP_EPS) ? 1 : 0;
          sh[j] = h;
        }
      }
    }
#endif
    if (chain && tid == 0) {
      double carry = (n > 0) ? d_abs(sd[n - 1]) : 0.0;  // |current d[j]|
      MGQP_UNROLL1
      for (int j = n - 1; j >= iq + 1; j--) {
#if MGQP_LARGE_SPLIT_CHAIN
        const double h = ln_distance_split(sd[j - 1], carry);
#else
        const double h = ln_distance(sd[j - 1], carry);
#endif
        if (d_abs(h) < MGQP_EPS) {
          skip[j] = 1;
          carry = d_abs(sd[j - 1]);
        } else {
          skip[j] = 0;
          sh[j] = h;
          carry = h;
        }
      }
    }
    blk_sync();
    // (2a) normalised cosine and the signed value each rotation leaves in d[j-1]
    MGQP_UNROLL1
    for (int j = iq + 1 + tid; j <= n - 1; j += T) {
      if (!skip[j]) {
        const double c = d_div(sd[j - 1], sh[j]);
        scc[j] = c;
        sxny[j] = (c < 0.0) ? -sh[j] : sh[j];  // new d[j-1] (parked in sxny until pass 2b)
      } else {
        sxny[j] = sd[j - 1];
      }
    }
    blk_sync();
    // (2b) sine, sign normalisation, s/(1+c).  Rotation j's incoming d[j] is what rotation j+1 left.
    MGQP_UNROLL1
    for (int j = iq + 1 + tid; j <= n - 1; j += T) {
      if (!skip[j]) {
        const double din = (j == n - 1) ? sd[n - 1] : sxny[j + 1];
        double s = d_div(din, sh[j]);
        double c = scc[j];
        if (c < 0.0) {
          c = -c;
          s = -s;
        }
        scc[j] = c;
        sss[j] = s;
        sh[j] = d_div(s, d_add(1.0, c));  // xny; h itself is no longer needed
      }
    }
    blk_sync();
    // d[iq] after the sweep: what rotation iq+1 left there, or the untouched entry if there was none
    const double dlast = (iq + 1 <= n - 1) ? sxny[iq + 1] : sd[iq];
    // (3) every thread streams its rows of J right to left through all rotations
    MGQP_UNROLL1
    for (int k = tid; k < n; k += T) {
      double* Jk = J + k;  // row k of J, element j at Jk[j * n]: coalesced across the threads
      double t2 = Jk[(size_t)(n - 1) * n];
      // The original entries J[k][j-1] do not depend on the rotations: they are fetched MGQP_LARGE_GB columns ahead of
      // their use.  (Written as one load and one store per trip, the loop waited a full memory round trip per
      // rotation -- ~300 of them per sweep: the stores to the same array keep the compiler from hoisting the loads.)
      constexpr int GB = MGQP_LARGE_GB;
      int j = n - 1;
      MGQP_UNROLL1
      for (; j - GB >= iq; j -= GB) {
        double t1v[GB];
#pragma unroll
        for (int u = 0; u < GB; u++) t1v[u] = Jk[(size_t)(j - 1 - u) * n];
#pragma unroll
        for (int u = 0; u < GB; u++) {
          const int jj = j - u;
          const double t1 = t1v[u];
          if (!skip[jj]) {
            const double n1 = d_add(d_mul(t1, scc[jj]), d_mul(t2, sss[jj]));
            const double n2 = d_sub(d_mul(sh[jj], d_add(t1, n1)), t2);
            Jk[(size_t)jj * n] = n2;
            if (!FMA) j_bad_mine = j_bad_mine || !(d_abs(n1) < INF) || !(d_abs(n2) < INF);
            t2 = n1;
          } else {
            Jk[(size_t)jj * n] = t2;
            t2 = t1;
          }
        }
      }
      MGQP_UNROLL1
      for (; j >= iq + 1; j--) {
        const double t1 = Jk[(size_t)(j - 1) * n];
        if (!skip[j]) {
          const double n1 = d_add(d_mul(t1, scc[j]), d_mul(t2, sss[j]));
          const double n2 = d_sub(d_mul(sh[j], d_add(t1, n1)), t2);
          Jk[(size_t)j * n] = n2;
          if (!FMA) j_bad_mine = j_bad_mine || !(d_abs(n1) < INF) || !(d_abs(n2) < INF);
          t2 = n1;
        } else {
          Jk[(size_t)j * n] = t2;
          t2 = t1;
        }
      }
      if (iq + 1 <= n - 1) Jk[(size_t)iq * n] = t2;
    }
    if (!FMA && j_finite) {  // (uniform: j_finite is a CTA-wide value)
      if (tid == 0) sredi[1] = 0;
      blk_sync();
      if (j_bad_mine) sredi[1] = 1;
      blk_sync();
      j_finite = sredi[1] == 0;
    }
    // column iq of R <- d[0..iq]
    MGQP_UNROLL1
    for (int i = tid; i < iq; i += T) R[(size_t)i * n + iq] = sd[i];
    if (tid == 0) R[(size_t)iq * n + iq] = dlast;
    iq++;
    blk_sync();
    if (d_abs(dlast) <= d_mul(MGQP_EPS, R_norm)) return false;
    R_norm = (R_norm < d_abs(dlast)) ? d_abs(dlast) : R_norm;
    return true;
  };

  // delete_constraint(l); the pending pair (A[iq], u[iq]) sits at index iq of A / su
  auto delete_constraint = [&](int l) {
    if (tid == 0) {
      int qq = 0;
      MGQP_UNROLL1
      for (int i = p; i < iq; i++)
        if (A[i] == l) {
          qq = i;
          break;
        }
      sredi[0] = qq;
    }
    blk_sync();
    const int qq = sredi[0];
    blk_sync();
    // shift columns qq+1.. of R one to the left (row-parallel), A and u likewise
    MGQP_UNROLL1
    for (int j = tid; j < n; j += T) {
      double* Rj = R + (size_t)j * n;
      MGQP_UNROLL1
      for (int i = qq; i < iq - 1; i++) Rj[i] = Rj[i + 1];
      if (j < iq) Rj[iq - 1] = 0.0;
    }
    if (tid == 0) {
      MGQP_UNROLL1
      for (int i = qq; i < iq - 1; i++) {
        A[i] = A[i + 1];
        su[i] = su[i + 1];
      }
      A[iq - 1] = A[iq];
      su[iq - 1] = su[iq];
      A[iq] = 0;
      su[iq] = 0.0;
    }
    iq--;
    blk_sync();
    if (iq == 0) return;
    MGQP_UNROLL1
    for (int j = qq; j < iq; j++) {
      if (tid == 0) {
        double cc = R[(size_t)j * n + j], ss = R[(size_t)(j + 1) * n + j];
        const double h = ln_distance(cc, ss);
        if (d_abs(h) < MGQP_EPS) {
          sbc[3] = 0.0;
        } else {
          cc = d_div(cc, h);
          ss = d_div(ss, h);
          R[(size_t)(j + 1) * n + j] = 0.0;
          if (cc < 0.0) {
            R[(size_t)j * n + j] = -h;
            cc = -cc;
            ss = -ss;
          } else {
            R[(size_t)j * n + j] = h;
          }
          sbc[0] = cc;
          sbc[1] = ss;
          sbc[2] = d_div(ss, d_add(1.0, cc));
          sbc[3] = 1.0;
        }
      }
      blk_sync();
      if (sbc[3] != 0.0) {
        const double cc = sbc[0], ss = sbc[1], xny = sbc[2];
        MGQP_UNROLL1
        for (int k = j + 1 + tid; k < iq; k += T) {
          const double t1 = R[(size_t)j * n + k], t2 = R[(size_t)(j + 1) * n + k];
          const double n1 = d_add(d_mul(t1, cc), d_mul(t2, ss));
          R[(size_t)j * n + k] = n1;
          R[(size_t)(j + 1) * n + k] = d_sub(d_mul(xny, d_add(t1, n1)), t2);
        }
        MGQP_UNROLL1
        for (int k = tid; k < n; k += T) {
          double* Jk = J + k;
          const double t1 = Jk[(size_t)j * n], t2 = Jk[(size_t)(j + 1) * n];
          const double n1 = d_add(d_mul(t1, cc), d_mul(t2, ss));
          const double n2 = d_sub(d_mul(xny, d_add(n1, t1)), t2);
          Jk[(size_t)j * n] = n1;
          Jk[(size_t)(j + 1) * n] = n2;
          if (!FMA) j_bad_mine = j_bad_mine || !(d_abs(n1) < INF) || !(d_abs(n2) < INF);
        }
      }
      blk_sync();
    }
    if (!FMA && j_finite) {
      if (tid == 0) sredi[1] = 0;
      blk_sync();
      if (j_bad_mine) sredi[1] = 1;
      blk_sync();
      j_finite = sredi[1] == 0;
    }
  };

#ifndef MGQP_EMULATE
  if (out.gate) {
    blk_sync();
    if (tid == 0) atomicSub(out.gate, 1u);
  }
#endif
  bool done = fail;

  // ---------------------------------------------------------------- equality constraints
  MGQP_UNROLL1
  for (int e = 0; e < p && !done; e++) {
    MGQP_UNROLL1
    for (int j = tid; j < n; j += T) snp[j] = MGQP_LDG(io.CE + (size_t)j * p + e);
    np_nnz = -1;
    blk_sync();
    direction();
    dot0(sz, sz, 0);
    dot0(snp, sx, 1);
    dot0(sz, snp, 2);
    blk_sync();
    const double zz = sbc[0], npx = sbc[1], znp = sbc[2];
    blk_sync();
    double t2 = 0.0;
    if (d_abs(zz) > MGQP_EPS) t2 = d_div(d_sub(-npx, MGQP_LDG(io.ce0 + e)), znp);
    MGQP_UNROLL1
    for (int k = tid; k < n; k += T) sx[k] = d_add(sx[k], d_mul(t2, sz[k]));
    MGQP_UNROLL1
    for (int k = tid; k < iq; k += T) su[k] = d_sub(su[k], d_mul(t2, sr[k]));
    if (tid == 0) {
      su[iq] = t2;
      A[e] = -e - 1;
    }
    f = d_add(f, d_mul(d_mul(0.5, d_mul(t2, t2)), znp));
    blk_sync();
    if (!add_constraint()) {
      status = 3;
      done = true;
    }
  }

  // ---------------------------------------------------------------- active-set iteration
  MGQP_UNROLL1
  for (int i = tid; i < m; i += T) act[i] = 0;
  // ELL form of CI: one pass over the dense block (thread = column, coalesced along the rows).  A zero
  // coefficient contributes +-0.0 to a sum that starts at +0.0 and can never become -0.0: leaving it out does not
  // change a bit of s_i as long as x is finite (0 * inf would be NaN); NaN coefficients count as nonzeros.
  bool any_sparse = false;
  if (!done && m > 0 && n <= 0xfffe) {
    MGQP_UNROLL1
    for (int i = tid; i < m; i += T) {
      int cnt = 0;
      MGQP_UNROLL4
      for (int j = 0; j < n; j++) {
        const double a = MGQP_LDG(io.CI + (size_t)j * m + i);
        if (a != 0.0) {
          if (cnt < EK) {
            ell_val[(size_t)cnt * m + i] = a;
            ell_row[(size_t)cnt * m + i] = (unsigned short)j;
          }
          cnt++;
        }
      }
      ell_nnz[i] = (cnt <= EK) ? (unsigned short)cnt : ELL_DENSE;
      any_sparse = any_sparse || (cnt <= EK && 4 * cnt <= n);
    }
    if (tid == 0) sredi[0] = 0;
    blk_sync();
    if (any_sparse) sredi[0] = 1;
    blk_sync();
    any_sparse = sredi[0] != 0;
  }
  blk_sync();
  const double thr = d_mul(d_mul(d_mul(d_mul((double)m, MGQP_EPS), c1), c2), 100.0);
  int ip = 0;
  double ss = 0.0, s_ip = 0.0;
  bool need_scan = true, l1 = true;

  MGQP_UNROLL1
  while (!done) {
    if (need_scan) {
      if (l1) {
        MGQP_UNROLL1
        for (int i = tid; i < m; i += T) excl[i] = 1;
        ss = 0.0;
        ip = 0;
        blk_sync();
      }
      double best = ss, psi = 0.0;
      int bidx = NONE;
      // the ELL form needs a finite x (see above); a poisoned iterate takes the dense scan
      bool sparse_scan = any_sparse;
      if (sparse_scan) {
        bool bad = false;
        MGQP_UNROLL1
        for (int j = tid; j < n; j += T) bad = bad || !(d_abs(sx[j]) < INF);
        if (tid == 0) sredi[0] = 0;
        blk_sync();
        if (bad) sredi[0] = 1;
        blk_sync();
        sparse_scan = sredi[0] == 0;
        blk_sync();
      }
      if (sparse_scan) {
        MGQP_UNROLL1
        for (int i0 = tid; i0 < m; i0 += 4 * T) {
          // four columns per thread at a time: four independent accumulation chains
          double acc[4] = {0.0, 0.0, 0.0, 0.0};
          int cnt[4], col[4];
          int kmax = 0;
#pragma unroll
          for (int c = 0; c < 4; c++) {
            col[c] = i0 + c * T;
            cnt[c] = (col[c] < m) ? (int)ell_nnz[col[c]] : 0;
            if (cnt[c] == (int)ELL_DENSE) {  // dense column: the reference's full sum
              double a = 0.0;
              MGQP_UNROLL4
              for (int j = 0; j < n; j++) a = mad<FMA>(MGQP_LDG(io.CI + (size_t)j * m + col[c]), sx[j], a);
              acc[c] = a;
              cnt[c] = 0;
            }
            kmax = cnt[c] > kmax ? cnt[c] : kmax;
          }
          // four nonzeros of each of the four columns per trip: 32 loads in flight (the lists live in L2 / DRAM)
          MGQP_UNROLL1
          for (int k0 = 0; k0 < kmax; k0 += 4) {
            double vv[4][4];
            int rr[4][4];
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
              for (int c = 0; c < 4; c++) {
                const bool on = k0 + u < cnt[c];
                const size_t at = (size_t)(k0 + u) * m + col[c];
                vv[u][c] = on ? ell_val[at] : 0.0;
                rr[u][c] = on ? (int)ell_row[at] : 0;
              }
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
              for (int c = 0; c < 4; c++)
                if (k0 + u < cnt[c]) acc[c] = mad<FMA>(vv[u][c], sx[rr[u][c]], acc[c]);
          }
#pragma unroll
          for (int c = 0; c < 4; c++) {
            const int i = col[c];
            if (i < m) {
              const double sum = d_add(acc[c], MGQP_LDG(io.ci0 + i));
              sg[i] = sum;
              if (sum < 0.0) psi = d_add(psi, sum);
              if (sum < best && !act[i] && excl[i]) {
                best = sum;
                bidx = i;
              }
            }
          }
        }
      } else
      MGQP_UNROLL1
      for (int i0 = tid; i0 < m; i0 += 4 * T) {
        // up to four columns per thread at a time: four independent accumulation chains
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        const int i1 = i0 + T, i2 = i0 + 2 * T, i3 = i0 + 3 * T;
        const bool v1 = i1 < m, v2 = i2 < m, v3 = i3 < m;
        MGQP_UNROLL4
        for (int j = 0; j < n; j++) {
          const double xj = sx[j];
          const double* row = io.CI + (size_t)j * m;
          a0 = mad<FMA>(MGQP_LDG(row + i0), xj, a0);
          if (v1) a1 = mad<FMA>(MGQP_LDG(row + i1), xj, a1);
          if (v2) a2 = mad<FMA>(MGQP_LDG(row + i2), xj, a2);
          if (v3) a3 = mad<FMA>(MGQP_LDG(row + i3), xj, a3);
        }
        const double acc[4] = {a0, a1, a2, a3};
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const int i = i0 + c * T;
          if (i < m) {
            const double sum = d_add(acc[c], MGQP_LDG(io.ci0 + i));
            sg[i] = sum;
            if (sum < 0.0) psi = d_add(psi, sum);
            if (sum < best && !act[i] && excl[i]) {
              best = sum;
              bidx = i;
            }
          }
        }
      }
      double bval;
      int bi;
      block_argmin(best, bidx, bval, bi);
      psi = block_sum(psi);
      if (l1) {
        if (!FMA && d_abs(psi) > d_mul(0.5, thr) && d_abs(psi) < d_mul(2.0, thr)) {
          // within 2x of the threshold: redo the sum in the reference's order (rare)
          if (tid == 0) {
            double q = 0.0;
            MGQP_UNROLL1
            for (int i = 0; i < m; i++) q = d_add(q, (sg[i] < 0.0) ? sg[i] : 0.0);
            sbc[0] = q;
          }
          blk_sync();
          psi = sbc[0];
          blk_sync();
        }
        if (d_abs(psi) <= thr) {
          done = true;
        } else {
          MGQP_UNROLL1
          for (int i = tid; i < iq; i += T) {
            suold[i] = su[i];
            Aold[i] = A[i];
          }
          MGQP_UNROLL1
          for (int i = tid; i < n; i += T) sxold[i] = sx[i];
        }
      }
      if (!done) {
        if (bi != NONE) {
          ss = bval;
          ip = bi;
          s_ip = bval;
        }
        if (ss >= 0.0) {
          done = true;
        } else {
          MGQP_UNROLL1
          for (int j = tid; j < n; j += T) snp[j] = MGQP_LDG(io.CI + (size_t)j * m + ip);
          np_nnz = -1;
          if (any_sparse && ell_nnz[ip] != ELL_DENSE) {
            np_nnz = (int)ell_nnz[ip];
            MGQP_UNROLL1
            for (int k = tid; k < np_nnz; k += T) {
              snzv[k] = ell_val[(size_t)k * m + ip];
              snzr[k] = (int)ell_row[(size_t)k * m + ip];
            }
          }
          if (tid == 0) {
            su[iq] = 0.0;
            A[iq] = ip;
          }
          need_scan = false;
        }
      }
      blk_sync();
    }
    if (done) break;

    steps++;
    if (steps > max_iter) {
      status = 4;
      break;
    }
    direction();

    // step lengths
    double ratio = INF;
    int lk = NONE;
    MGQP_UNROLL1
    for (int k = p + tid; k < iq; k += T) {
      if (sr[k] > 0.0) {
        const double q = d_div(su[k], sr[k]);
        if (q < ratio) {  // ascending k within a thread: first minimum wins
          ratio = q;
          lk = k;
        }
      }
    }
    double t1;
    int lkk;
    block_argmin(ratio, lk, t1, lkk);
    const int l = (lkk == NONE) ? 0 : A[lkk];
    dot0(sz, sz, 0);
    dot0(sz, snp, 1);
    blk_sync();
    const double zz = sbc[0], znp = sbc[1];
    const double uq = su[iq];
    blk_sync();
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
      // dual step only
      MGQP_UNROLL1
      for (int k = tid; k < iq; k += T) su[k] = d_sub(su[k], d_mul(t, sr[k]));
      if (tid == 0) {
        su[iq] = d_add(uq, t);
        act[l] = 0;
      }
      blk_sync();
      delete_constraint(l);
      continue;
    }
    MGQP_UNROLL1
    for (int k = tid; k < n; k += T) sx[k] = d_add(sx[k], d_mul(t, sz[k]));
    f = d_add(f, d_mul(d_mul(t, znp), d_add(d_mul(0.5, t), uq)));
    MGQP_UNROLL1
    for (int k = tid; k < iq; k += T) su[k] = d_sub(su[k], d_mul(t, sr[k]));
    if (tid == 0) su[iq] = d_add(uq, t);
    blk_sync();
    if (d_abs(d_sub(t, t2)) < MGQP_EPS) {
      if (!add_constraint()) {
        if (tid == 0) excl[ip] = 0;
        blk_sync();
        delete_constraint(ip);
        MGQP_UNROLL1
        for (int i = tid; i < m; i += T) act[i] = 0;
        blk_sync();
        MGQP_UNROLL1
        for (int i = p + tid; i < iq; i += T) {
          A[i] = Aold[i];
          su[i] = suold[i];
          act[Aold[i]] = 1;
        }
        MGQP_UNROLL1
        for (int i = tid; i < n; i += T) sx[i] = sxold[i];
        need_scan = true;
        l1 = false;
      } else {
        if (tid == 0) act[ip] = 1;
        need_scan = true;
        l1 = true;
      }
      blk_sync();
      continue;
    }
    // partial step
    if (tid == 0) act[l] = 0;
    blk_sync();
    delete_constraint(l);
    dot0(snp, sx, 0);
    blk_sync();
    s_ip = d_add(sbc[0], MGQP_LDG(io.ci0 + ip));
    blk_sync();
  }

  // ---------------------------------------------------------------- results
  blk_sync();
  MGQP_UNROLL1
  for (int i = tid; i < n; i += T) out.x[i] = fail ? 0.0 : sx[i];
  if (tid == 0) {
    *out.f = f;
    *out.status = status;
    if (out.iq) *out.iq = iq;
    if (out.iters) *out.iters = steps;
  }
  if (out.A) {
    MGQP_UNROLL1
    for (int i = tid; i < p + m; i += T) out.A[i] = (i < iq) ? A[i] : NONE;
  }
  blk_sync();
}

}  // namespace mgqp
