// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// C-ABI shim around the reference's *shipped* solver binary
// (/root/reference/lib/QuadProgpp/libquadprog.a, GCC 5.4 -O0, source absent), built against the
// reference's own headers where they lie (/root/reference/include/QuadProgpp/{QuadProg++.hh,Array.hh});
// nothing from /root/reference is copied into this repository.  The output goes to oracle/_ref/
// (git-ignored, shipped to the GPU box by gpurun).  See oracle/Makefile.
//
// Wraps: double solve_quadprog(Matrix<double>& G, Vector<double>& g0, const Matrix<double>& CE,
//        const Vector<double>& ce0, const Matrix<double>& CI, const Vector<double>& ci0,
//        Vector<double>& x)                         -- include/QuadProgpp/QuadProg++.hh:69-72
// as called by the reference adapter at src/mgqp.cpp:708 and :725.
#include <cstring>
#include <cstdio>
#include <cmath>
#include <limits>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#include <atomic>
#include <unistd.h>
#include <fcntl.h>

#include "QuadProg++.hh"

namespace {

// The shipped binary prints the offending matrix to std::cout before throwing on a non-PD G
// (SURVEY.md section 5).  Silence fd 1 around that so bulk runs do not flood the log.
struct StdoutMute {
  int saved = -1;
  explicit StdoutMute(bool on) {
    if (!on) return;
    fflush(stdout);
    saved = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    if (nul >= 0) { dup2(nul, 1); close(nul); }
  }
  ~StdoutMute() {
    if (saved >= 0) { fflush(stdout); std::cout.flush(); dup2(saved, 1); close(saved); }
  }
};

int solve_one(int n, int p, int m, const double* G, const double* g0, const double* CE,
              const double* ce0, const double* CI, const double* ci0, double* x, double* f,
              double* G_out, char* msg, int msg_len) {
  ArrayHH::Matrix<double> mG(n, n), mCE(n, p), mCI(n, m);
  ArrayHH::Vector<double> vg0(n), vce0(p), vci0(m), vx(n);
  for (int i = 0; i < n; i++) {
    for (int j = 0; j < n; j++) mG[i][j] = G[(size_t)i * n + j];
    for (int j = 0; j < p; j++) mCE[i][j] = CE[(size_t)i * p + j];
    for (int j = 0; j < m; j++) mCI[i][j] = CI[(size_t)i * m + j];
    vg0[i] = g0[i];
    vx[i] = 0.0;
  }
  for (int j = 0; j < p; j++) vce0[j] = ce0[j];
  for (int j = 0; j < m; j++) vci0[j] = ci0[j];
  int status = 0;
  double cost = std::numeric_limits<double>::quiet_NaN();
  try {
    cost = solve_quadprog(mG, vg0, mCE, vce0, mCI, vci0, vx);
    if (cost == std::numeric_limits<double>::infinity()) status = 1;
  } catch (const std::runtime_error& e) {
    status = 3;  // "Constraints are linearly dependent"
    if (msg && msg_len > 0) { std::strncpy(msg, e.what(), msg_len - 1); msg[msg_len - 1] = 0; }
  } catch (const std::logic_error& e) {
    status = 2;  // non-PD G ("Error in cholesky decomposition, sum: ...") or dimension mismatch
    if (msg && msg_len > 0) { std::strncpy(msg, e.what(), msg_len - 1); msg[msg_len - 1] = 0; }
  }
  *f = cost;
  for (int i = 0; i < n; i++) x[i] = vx[i];
  if (G_out)
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) G_out[(size_t)i * n + j] = mG[i][j];
  return status;
}

}  // namespace

extern "C" {

// One QP.  Layout identical to QuadProg++: G n*n, CE n*p, CI n*m row-major (constraint = column).
// status: 0 ok, 1 infeasible (+inf returned), 2 logic_error, 3 runtime_error.  G_out (optional)
// receives G as the binary leaves it (its Cholesky factor, QuadProg++.hh:42-45).
int ref_solve_quadprog(int n, int p, int m, const double* G, const double* g0, const double* CE,
                       const double* ce0, const double* CI, const double* ci0, double* x, double* f,
                       double* G_out, char* msg, int msg_len) {
  StdoutMute mute(true);
  return solve_one(n, p, m, G, g0, CE, ce0, CI, ci0, x, f, G_out, msg, msg_len);
}

// Dimension-mismatch probe: lets a test hand the binary a CE with the wrong row count
// (SURVEY.md 8c K7).  Returns the exception text.
int ref_dim_mismatch_message(int n, int ce_rows, char* msg, int msg_len) {
  ArrayHH::Matrix<double> G(n, n), CE(ce_rows, 1), CI(n, 0);
  ArrayHH::Vector<double> g0(n), ce0(1), ci0(0u), x(n);
  for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) G[i][j] = (i == j); g0[i] = 0; }
  for (int i = 0; i < ce_rows; i++) CE[i][0] = 1.0;
  ce0[0] = 0.0;
  try {
    solve_quadprog(G, g0, CE, ce0, CI, ci0, x);
  } catch (const std::logic_error& e) {
    std::strncpy(msg, e.what(), msg_len - 1); msg[msg_len - 1] = 0;
    return 2;
  }
  return 0;
}

// Batch over contiguous SoA buffers ([batch][rows][cols]); `threads` host threads, dynamic chunks.
// The shipped binary is re-entrant (no mutable globals; SURVEY.md 8b "Threading").
// Returns seconds spent inside the solve loop (steady clock) through *seconds.
int ref_solve_quadprog_batch(int n, int p, int m, long long batch, const double* G, const double* g0,
                             const double* CE, const double* ce0, const double* CI, const double* ci0,
                             double* x, double* f, int* status, int threads) {
  StdoutMute mute(true);
  if (threads < 1) threads = 1;
  std::atomic<long long> next(0);
  const long long chunk = 64;
  auto work = [&]() {
    for (;;) {
      long long b0 = next.fetch_add(chunk);
      if (b0 >= batch) break;
      long long b1 = b0 + chunk < batch ? b0 + chunk : batch;
      for (long long b = b0; b < b1; b++) {
        status[b] = solve_one(n, p, m, G + b * n * n, g0 + b * n, CE + b * n * p, ce0 + b * p,
                              CI + b * n * m, ci0 + b * m, x + b * n, f + b, nullptr, nullptr, 0);
      }
    }
  };
  if (threads == 1) { work(); return 0; }
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) pool.emplace_back(work);
  for (auto& t : pool) t.join();
  return 0;
}

}  // extern "C"
