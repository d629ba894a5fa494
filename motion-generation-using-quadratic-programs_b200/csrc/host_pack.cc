// Host side of the column-compressed constraint transport (mgqp_batch::CI_val / CI_row in
// include/mgqp_b200.h): turns the dense n x m inequality block the reference's caller builds
// (src/mgqp.cpp:1080-1128: limitsMatrix = [-I; +I], one nonzero per column) into m (value, row)
// pairs per QP, so that the PCIe link carries m*9 instead of n*m*8 bytes.  This is a transport
// codec, not a solver: no arithmetic of the QP is done here, and a chunk in which any column has a
// second nonzero is shipped dense instead.
//
// Plain C++ compiled by the host compiler (not nvcc's front end): the inner loop has an AVX2 version
// picked at run time.
#include <stdint.h>
#include <string.h>

#include <sched.h>
#include <stdlib.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/mgqp_b200.h"

namespace {

// One QP (m <= 64: a row's nonzero pattern fits one 64-bit mask).  Each row of CI becomes a bit mask of
// its nonzero entries (+-0.0 are the zeros, NaN counts as nonzero); a column with two nonzeros shows up as
// a bit common to two rows.  The ~m nonzeros are then scattered into (val, row); an all-zero column gets
// (+0.0, row 0).  pack_one_* return 1 if some column has more than one nonzero.
inline void scatter_row(uint64_t mask, int j, const double* r, double* val, uint8_t* row) {
  while (mask) {
    const int i = __builtin_ctzll(mask);
    mask &= mask - 1;
    val[i] = r[i];
    row[i] = (uint8_t)j;
  }
}

int pack_one_generic(int n, int m, const double* __restrict__ ci, double* __restrict__ val, uint8_t* __restrict__ row) {
  for (int i = 0; i < m; i++) {
    val[i] = 0.0;
    row[i] = 0;
  }
  uint64_t seen = 0, dup = 0;
  for (int j = 0; j < n; j++) {
    const double* r = ci + (int64_t)j * m;
    uint64_t mask = 0;
    for (int i = 0; i < m; i++) {
      uint64_t bits;
      memcpy(&bits, &r[i], 8);
      mask |= (uint64_t)((bits << 1) != 0) << i;
    }
    dup |= seen & mask;
    seen |= mask;
    scatter_row(mask, j, r, val, row);
  }
  return dup != 0;
}

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
__attribute__((target("avx2"))) int pack_one_avx2(int n, int m, const double* __restrict__ ci, double* __restrict__ val,
                                                  uint8_t* __restrict__ row) {
  const __m256d zero = _mm256_setzero_pd();
  int i = 0;
  for (; i + 4 <= m; i += 4) _mm256_storeu_pd(val + i, zero);
  for (; i < m; i++) val[i] = 0.0;
  memset(row, 0, (size_t)m);
  uint64_t seen = 0, dup = 0;
  for (int j = 0; j < n; j++) {
    const double* r = ci + (int64_t)j * m;
    uint64_t mask = 0;
    int k = 0;
    for (; k + 4 <= m; k += 4) {
      const __m256d v = _mm256_loadu_pd(r + k);
      mask |= (uint64_t)_mm256_movemask_pd(_mm256_cmp_pd(v, zero, _CMP_NEQ_UQ)) << k;
    }
    for (; k < m; k++) {
      uint64_t bits;
      memcpy(&bits, &r[k], 8);
      mask |= (uint64_t)((bits << 1) != 0) << k;
    }
    dup |= seen & mask;
    seen |= mask;
    scatter_row(mask, j, r, val, row);
  }
  return dup != 0;
}
#define MGQP_HAVE_AVX2_PACK 1

// AVX-512: eight columns per vector, no data-dependent branch.  Per row: one load, one test ((bits << 1) != 0 as
// an integer test, so -0.0 is a zero and NaN a nonzero whatever MXCSR says), two mask-register operations for
// seen / dup, two masked moves that keep the row's value and index where it is nonzero.  The scalar scatter of the
// AVX2 version (a count-trailing-zeros loop per row, one mispredicted exit each) was what bounded the packer:
// profiles/r02_notes.md.
__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq"))) int pack_one_avx512(int n, int m, const double* __restrict__ ci,
                                                                            double* __restrict__ val,
                                                                            uint8_t* __restrict__ row) {
  const __m512i absmask = _mm512_set1_epi64(0x7fffffffffffffffLL);
  unsigned dup = 0;
  for (int k = 0; k < m; k += 8) {
    const int left = m - k;
    const __mmask8 km = left >= 8 ? (__mmask8)0xff : (__mmask8)((1u << left) - 1u);
    __m512i vb = _mm512_setzero_si512(), rv = _mm512_setzero_si512();
    __mmask8 seen = 0;
    const double* r = ci + k;
    for (int j = 0; j < n; j++, r += m) {
      const __m512i v = _mm512_maskz_loadu_epi64(km, r);
      const __mmask8 nz = _mm512_test_epi64_mask(v, absmask);
      dup |= (unsigned)(seen & nz);
      seen |= nz;
      vb = _mm512_mask_mov_epi64(vb, nz, v);
      rv = _mm512_mask_mov_epi64(rv, nz, _mm512_set1_epi64(j));
    }
    _mm512_mask_storeu_epi64(val + k, km, vb);
    _mm512_mask_cvtepi64_storeu_epi8(row + k, km, rv);
  }
  return dup != 0;
}
#define MGQP_HAVE_AVX512_PACK 1
#endif

// ---- unit-coefficient columns: ONE byte per column --------------------------------------------------------------
// Every inequality the reference's caller builds is +-1 times a unit vector (limitsMatrix = [-I; +I],
// src/mgqp.cpp:1111-1112).  For such a chunk the (value, row) pair collapses into a code byte: bits 0-5 the row,
// bit 7 set for -1.0, 0x40 for an all-zero column; the device expands it back into (value, row)
// (unit_code_expand_kernel in mgqp_b200.cu).  unit_one_* return 1 if some column has two nonzeros or a nonzero that
// is not exactly +-1.0 -- the chunk then goes through pack_one_* instead.
constexpr uint8_t kUnitEmpty = 0x40;

int unit_one_generic(int n, int m, const double* __restrict__ ci, uint8_t* __restrict__ code) {
  for (int i = 0; i < m; i++) code[i] = kUnitEmpty;
  int bad = 0;
  for (int j = 0; j < n; j++) {
    const double* r = ci + (int64_t)j * m;
    for (int i = 0; i < m; i++) {
      uint64_t bits;
      memcpy(&bits, &r[i], 8);
      if ((bits << 1) == 0) continue;
      if (code[i] != kUnitEmpty || (bits << 1) != (0x3ff0000000000000ull << 1)) bad = 1;
      code[i] = (uint8_t)(j | ((bits >> 56) & 0x80));
    }
  }
  return bad;
}

#ifdef MGQP_HAVE_AVX512_PACK
__attribute__((target("avx512f,avx512bw,avx512vl,avx512dq"))) int unit_one_avx512(int n, int m, const double* __restrict__ ci,
                                                                            uint8_t* __restrict__ code) {
  const __m512i absmask = _mm512_set1_epi64(0x7fffffffffffffffLL);
  const __m512i one = _mm512_set1_epi64(0x3ff0000000000000LL);
  const __m512i bit7 = _mm512_set1_epi64(0x80);
  unsigned bad = 0;
  for (int k = 0; k < m; k += 8) {
    const int left = m - k;
    const __mmask8 km = left >= 8 ? (__mmask8)0xff : (__mmask8)((1u << left) - 1u);
    __m512i cv = _mm512_set1_epi64(kUnitEmpty);
    __mmask8 seen = 0;
    const double* r = ci + k;
    for (int j = 0; j < n; j++, r += m) {
      const __m512i v = _mm512_maskz_loadu_epi64(km, r);
      const __m512i a = _mm512_and_si512(v, absmask);
      const __mmask8 nz = _mm512_test_epi64_mask(a, a);
      const __mmask8 unit = _mm512_cmpeq_epi64_mask(a, one);
      bad |= (unsigned)(seen & nz) | (unsigned)(nz & (__mmask8)~unit);
      seen |= nz;
      const __m512i c = _mm512_or_si512(_mm512_and_si512(_mm512_srli_epi64(v, 56), bit7), _mm512_set1_epi64(j));
      cv = _mm512_mask_mov_epi64(cv, nz, c);
    }
    _mm512_mask_cvtepi64_storeu_epi8(code + k, km, cv);
  }
  return bad != 0;
}
#endif

typedef int (*unit_fn)(int, int, const double*, uint8_t*);
unit_fn pick_unit() {
  const char* force = getenv("MGQP_HOST_PACK_ISA");
#ifdef MGQP_HAVE_AVX512_PACK
  if ((!force || !strcmp(force, "avx512")) && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
      __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512dq"))
    return unit_one_avx512;
#endif
  return unit_one_generic;
}

typedef int (*pack_fn)(int, int, const double*, double*, uint8_t*);
pack_fn pick_pack() {
  const char* force = getenv("MGQP_HOST_PACK_ISA");  // tests: "generic" / "avx2" / "avx512"
#ifdef MGQP_HAVE_AVX512_PACK
  if ((!force || !strcmp(force, "avx512")) && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") &&
      __builtin_cpu_supports("avx512vl") && __builtin_cpu_supports("avx512dq"))
    return pack_one_avx512;
#endif
#ifdef MGQP_HAVE_AVX2_PACK
  if ((!force || strcmp(force, "generic")) && __builtin_cpu_supports("avx2")) return pack_one_avx2;
#endif
  return pack_one_generic;
}

// A small persistent pool: the host pipeline packs one chunk (tens of thousands of QPs, ~2 ms of
// work) at a time, far too often to create threads for each.  One job at a time; the caller works too.
class Pool {
 public:
  explicit Pool(int workers) {
    for (int i = 0; i < workers; i++) th_.emplace_back([this]() { loop(); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int workers() const { return (int)th_.size(); }
  // run job() on `extra` pool threads and on the calling thread
  void run(const std::function<void()>& job, int extra) {
    std::lock_guard<std::mutex> one(call_mu_);
    // the pool grows with the largest request it has seen (MGQP_HOST_THREADS may be raised between calls: a rank
    // that packed with its share of the cores later drives every device of the node on its own)
    while ((int)th_.size() < extra && (int)th_.size() < 63) th_.emplace_back([this]() { loop(); });
    {
      std::lock_guard<std::mutex> lk(mu_);
      job_ = &job;
      want_ = extra;
      pending_ = extra;
      gen_++;
    }
    cv_.notify_all();
    job();
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this]() { return pending_ == 0; });
    job_ = nullptr;
  }

 private:
  void loop() {
    long seen = 0;
    for (;;) {
      const std::function<void()>* job = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&]() { return stop_ || (gen_ != seen && want_ > 0); });
        if (stop_) return;
        seen = gen_;
        want_--;
        job = job_;
      }
      (*job)();
      {
        std::lock_guard<std::mutex> lk(mu_);
        pending_--;
      }
      done_.notify_all();
    }
  }
  std::mutex mu_, call_mu_;
  std::condition_variable cv_, done_;
  std::vector<std::thread> th_;
  const std::function<void()>* job_ = nullptr;
  long gen_ = 0;
  int want_ = 0, pending_ = 0;
  bool stop_ = false;
};

int default_threads() {
  if (const char* e = getenv("MGQP_HOST_THREADS")) {
    const int t = atoi(e);
    if (t >= 1) return t > 64 ? 64 : t;
  }
  int t = (int)std::thread::hardware_concurrency();
#if defined(__linux__)
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof set, &set) == 0) {  // a container may own fewer cores than the machine has
    const int a = CPU_COUNT(&set);
    if (a >= 1 && a < t) t = a;
  }
#endif
  if (t < 1) t = 1;
  return t > 64 ? 64 : t;
}

int pack_prefetch_ahead() {
  static const int a = []() {
    const char* e = getenv("MGQP_HOST_PACK_PREFETCH");
    return e ? atoi(e) : 2;  // QPs ahead; measured on 1-4 threads: 0 -> 2 = +40 % (profiles/r02_notes.md)
  }();
  return a;
}

Pool& pool() {
  static Pool p(default_threads() - 1);
  return p;
}

}  // namespace

extern "C" int mgqp_host_pack_threads(void) { return default_threads(); }

// 1 iff every one of `count` blocks of `size` doubles (element stride `stride` between blocks) is bit for
// bit the first one: the reference's own caller passes G = I and g0 = 0 for every QP (src/mgqp.cpp:672,
// :706), and such a field can cross the link once per chunk instead of once per QP.
extern "C" int mgqp_blocks_all_equal(const double* base, int64_t stride, int64_t size, int64_t count, int threads) {
  if (count <= 1 || size <= 0) return 1;
  if (threads <= 0) threads = default_threads();
  const int64_t block = 4096;
  const int64_t blocks = (count + block - 1) / block;
  if (threads > blocks) threads = (int)blocks;
  std::atomic<int64_t> next(0);
  std::atomic<int> differ(0);
  const size_t bytes = (size_t)size * 8;
  const std::function<void()> work = [&]() {
    for (;;) {
      const int64_t b = next.fetch_add(1);
      if (b >= blocks || differ.load(std::memory_order_relaxed)) break;
      const int64_t q1 = (b + 1) * block < count ? (b + 1) * block : count;
      for (int64_t q = b * block; q < q1; q++)
        if (memcmp(base, base + q * stride, bytes) != 0) {
          differ.store(1, std::memory_order_relaxed);
          break;
        }
    }
  };
  if (threads <= 1) work();
  else pool().run(work, threads - 1);
  return differ.load() ? 0 : 1;
}

extern "C" int mgqp_pack_bound_columns(int32_t n, int32_t m, int64_t count, const double* CI, int64_t sCI, double* val,
                                       int64_t sv, uint8_t* row, int64_t sr, int threads) {
  if (n < 1 || n > 255 || m < 0 || m > 64 || count < 0) return 0;
  if (count == 0 || m == 0) return 1;
  static const pack_fn pack_one = pick_pack();
  if (threads <= 0) threads = default_threads();
  const int64_t block = 512;
  const int64_t blocks = (count + block - 1) / block;
  if (threads > blocks) threads = (int)blocks;
  std::atomic<int64_t> next(0);
  std::atomic<int> dense(0);
  const std::function<void()> work = [&]() {
    for (;;) {
      const int64_t b = next.fetch_add(1);
      if (b >= blocks || dense.load(std::memory_order_relaxed)) break;
      const int64_t q1 = (b + 1) * block < count ? (b + 1) * block : count;
      int bad = 0;
      const int64_t qp_bytes = (int64_t)n * m * 8;
      const int ahead = pack_prefetch_ahead();
      for (int64_t q = b * block; q < q1; q++) {
        if (ahead > 0 && q + ahead < count) {  // the packer is one sequential stream per thread: keep more lines in flight
          const char* nx = reinterpret_cast<const char*>(CI + (q + ahead) * sCI);
          for (int64_t o = 0; o < qp_bytes; o += 64) __builtin_prefetch(nx + o, 0, 3);
        }
        bad |= pack_one(n, m, CI + q * sCI, val + q * sv, row + q * sr);
      }
      if (bad) dense.store(1, std::memory_order_relaxed);
    }
  };
  if (threads <= 1) work();
  else pool().run(work, threads - 1);
  return dense.load() ? 0 : 1;
}

// Unit-coefficient form: one code byte per column (see unit_one_generic).  Returns 1 if every column of every QP is
// empty or exactly +-1.0 times a unit vector, 0 otherwise (`code` is then unspecified).
extern "C" int mgqp_pack_unit_columns(int32_t n, int32_t m, int64_t count, const double* CI, int64_t sCI, uint8_t* code,
                                      int64_t sc, int threads) {
  if (n < 1 || n > 64 || m < 0 || m > 64 || count < 0) return 0;
  if (count == 0 || m == 0) return 1;
  static const unit_fn unit_one = pick_unit();
  if (threads <= 0) threads = default_threads();
  const int64_t block = 512;
  const int64_t blocks = (count + block - 1) / block;
  if (threads > blocks) threads = (int)blocks;
  std::atomic<int64_t> next(0);
  std::atomic<int> refused(0);
  const std::function<void()> work = [&]() {
    for (;;) {
      const int64_t b = next.fetch_add(1);
      if (b >= blocks || refused.load(std::memory_order_relaxed)) break;
      const int64_t q1 = (b + 1) * block < count ? (b + 1) * block : count;
      const int64_t qp_bytes = (int64_t)n * m * 8;
      const int ahead = pack_prefetch_ahead();
      int bad = 0;
      for (int64_t q = b * block; q < q1; q++) {
        if (ahead > 0 && q + ahead < count) {
          const char* nx = reinterpret_cast<const char*>(CI + (q + ahead) * sCI);
          for (int64_t o = 0; o < qp_bytes; o += 64) __builtin_prefetch(nx + o, 0, 3);
        }
        bad |= unit_one(n, m, CI + q * sCI, code + q * sc);
      }
      if (bad) refused.store(1, std::memory_order_relaxed);
    }
  };
  if (threads <= 1) work();
  else pool().run(work, threads - 1);
  return refused.load() ? 0 : 1;
}

// 1 iff the unit-code packer has a vector version on this CPU (the scalar one is slower than the (value, row) packer)
extern "C" int mgqp_pack_unit_is_fast(void) {
  static const int fast = pick_unit() != unit_one_generic;
  return fast;
}
