// host_pack.cu -- host-side input glue of the drop-in face (no device code): FeatureGroup.construct_inputs hands Keras one
// array per feature (features/group.py:218-248: (N,1) ids, (N,seq_len) histories, (N,dim) dense values); the fused lookup
// wants ONE packed int32 id matrix and ONE fp32 dense matrix per batch.  These two entry points gather the
// rows [row_start, row_start+rows) of every column array into a (pinned) destination matrix with a small persistent thread
// pool, so that Model.fit can assemble batch t+1 on the host while batch t trains on the GPU.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <type_traits>
#include <vector>

#if defined(__x86_64__)
#include <cpuid.h>
#include <immintrin.h>
#endif

#include "common.cuh"

namespace {

class Pool {
 public:
  explicit Pool(int n) : stop_(false), pending_(0), gen_(0) {
    for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { loop(i); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> l(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return (int)workers_.size(); }
  // run fn(worker) on workers [0, n) and wait
  void run(int n, const std::function<void(int)>& fn) {
    std::unique_lock<std::mutex> l(mu_);
    fn_ = &fn;
    active_ = n;
    pending_ = n;
    ++gen_;
    cv_.notify_all();
    done_.wait(l, [this] { return pending_ == 0; });
  }

 private:
  void loop(int id) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)>* fn = nullptr;
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_.wait(l, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
        if (id >= active_) continue;
        fn = fn_;
      }
      (*fn)(id);
      {
        std::lock_guard<std::mutex> l(mu_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  bool stop_;
  int pending_, active_ = 0;
  uint64_t gen_;
};

Pool& pool() {
  static Pool p((int)std::max(1u, std::min(16u, std::thread::hardware_concurrency())));
  return p;
}
std::mutex g_pack_mu;  // one packing job at a time (the pool is shared)

// Packed batches are read next by the GPU's copy engine over PCIe.  Lines left dirty in the packing cores' private caches make that
// DMA snoop them out one by one (measured: the last batches of a fit, which no later packing evicts, slowed the step they
// overlapped by up to 0.8 ms).  Writing every finished block back (clwb / clflushopt) leaves the data in memory for the DMA.
#if defined(__x86_64__)
bool have_avx2() {
  static const bool ok = __builtin_cpu_supports("avx2");
  return ok;
}
#endif
std::atomic<int> g_writeback{1};
int cache_writeback_kind() {  // 2 = clwb, 1 = clflushopt, 0 = neither
  static const int kind = [] {
#if defined(__x86_64__)
    unsigned a = 0, b = 0, c = 0, d = 0;
    if (__get_cpuid_count(7, 0, &a, &b, &c, &d)) return (b & (1u << 24)) ? 2 : ((b & (1u << 23)) ? 1 : 0);
#endif
    return 0;
  }();
  return kind;
}
inline void writeback_range(const void* p, size_t bytes) {
#if defined(__x86_64__)
  const int kind = cache_writeback_kind();
  if (!kind || !g_writeback.load(std::memory_order_relaxed)) return;
  const char* a = (const char*)((uintptr_t)p & ~(uintptr_t)63);
  const char* e = (const char*)p + bytes;
  if (kind == 2) for (; a < e; a += 64) asm volatile(".byte 0x66; xsaveopt %0" : "+m"(*(volatile char*)a));  // clwb
  else for (; a < e; a += 64) asm volatile(".byte 0x66; clflush %0" : "+m"(*(volatile char*)a));             // clflushopt
#else
  (void)p; (void)bytes;
#endif
}

#if defined(__x86_64__)
// 8 columns x 8 rows of 32-bit words -> 8 destination rows of 8 words each (AVX2 in-register transpose).  The scalar walk spends
// ~6 ns per element on 4-byte strided stores; this writes 32 contiguous bytes per store.
__attribute__((target("avx2"))) inline void transpose8x8_store(const uint32_t* const c[8], int64_t r, uint32_t* d, int64_t dst_ld) {
  __m256i v0 = _mm256_loadu_si256((const __m256i*)(c[0] + r)), v1 = _mm256_loadu_si256((const __m256i*)(c[1] + r));
  __m256i v2 = _mm256_loadu_si256((const __m256i*)(c[2] + r)), v3 = _mm256_loadu_si256((const __m256i*)(c[3] + r));
  __m256i v4 = _mm256_loadu_si256((const __m256i*)(c[4] + r)), v5 = _mm256_loadu_si256((const __m256i*)(c[5] + r));
  __m256i v6 = _mm256_loadu_si256((const __m256i*)(c[6] + r)), v7 = _mm256_loadu_si256((const __m256i*)(c[7] + r));
  __m256i t0 = _mm256_unpacklo_epi32(v0, v1), t1 = _mm256_unpackhi_epi32(v0, v1);
  __m256i t2 = _mm256_unpacklo_epi32(v2, v3), t3 = _mm256_unpackhi_epi32(v2, v3);
  __m256i t4 = _mm256_unpacklo_epi32(v4, v5), t5 = _mm256_unpackhi_epi32(v4, v5);
  __m256i t6 = _mm256_unpacklo_epi32(v6, v7), t7 = _mm256_unpackhi_epi32(v6, v7);
  __m256i u0 = _mm256_unpacklo_epi64(t0, t2), u1 = _mm256_unpackhi_epi64(t0, t2);
  __m256i u2 = _mm256_unpacklo_epi64(t1, t3), u3 = _mm256_unpackhi_epi64(t1, t3);
  __m256i u4 = _mm256_unpacklo_epi64(t4, t6), u5 = _mm256_unpackhi_epi64(t4, t6);
  __m256i u6 = _mm256_unpacklo_epi64(t5, t7), u7 = _mm256_unpackhi_epi64(t5, t7);
  _mm256_storeu_si256((__m256i*)(d + 0 * dst_ld), _mm256_permute2x128_si256(u0, u4, 0x20));
  _mm256_storeu_si256((__m256i*)(d + 1 * dst_ld), _mm256_permute2x128_si256(u1, u5, 0x20));
  _mm256_storeu_si256((__m256i*)(d + 2 * dst_ld), _mm256_permute2x128_si256(u2, u6, 0x20));
  _mm256_storeu_si256((__m256i*)(d + 3 * dst_ld), _mm256_permute2x128_si256(u3, u7, 0x20));
  _mm256_storeu_si256((__m256i*)(d + 4 * dst_ld), _mm256_permute2x128_si256(u0, u4, 0x31));
  _mm256_storeu_si256((__m256i*)(d + 5 * dst_ld), _mm256_permute2x128_si256(u1, u5, 0x31));
  _mm256_storeu_si256((__m256i*)(d + 6 * dst_ld), _mm256_permute2x128_si256(u2, u6, 0x31));
  _mm256_storeu_si256((__m256i*)(d + 7 * dst_ld), _mm256_permute2x128_si256(u3, u7, 0x31));
}

// rows [a, b) of n_cols >= 8 unit-stride 32-bit columns -> dst (row-major).  Column groups start at 0, 8, ... and the last one
// at n_cols - 8 (it re-writes up to 7 columns of its neighbour with the same values).
__attribute__((target("avx2"))) void pack_rows_avx2(const uint32_t* const* cols, int n_cols, int64_t a, int64_t b, int64_t row_start,
                                                    uint32_t* dst, int64_t dst_ld) {
  const int64_t b8 = a + (b - a) / 8 * 8;
  for (int64_t r = a; r < b8; r += 8) {
    uint32_t* d = dst + (r - row_start) * dst_ld;
    for (int c0 = 0; c0 < n_cols; c0 += 8) {
      const int c = c0 + 8 <= n_cols ? c0 : n_cols - 8;
      transpose8x8_store(cols + c, r, d + c, dst_ld);
    }
  }
  for (int64_t r = b8; r < b; ++r) {
    uint32_t* d = dst + (r - row_start) * dst_ld;
    for (int c = 0; c < n_cols; ++c) d[c] = cols[c][r];
  }
}
#endif

#if defined(__x86_64__)
uint32_t* thread_scratch(size_t words) {
  static thread_local uint32_t* buf = nullptr;
  static thread_local size_t cap = 0;
  if (cap < words) {
    free(buf);
    buf = static_cast<uint32_t*>(aligned_alloc(64, (words * 4 + 63) / 64 * 64));
    cap = buf ? words : 0;
  }
  return buf;
}
// dst 64-byte aligned, src 64-byte aligned scratch: 32-byte non-temporal stores for the body, plain stores + write-back for the tail
__attribute__((target("avx2"))) void stream_copy(uint32_t* dst, const uint32_t* src, size_t bytes) {
  const size_t body = bytes / 32 * 32;
  for (size_t o = 0; o < body; o += 32)
    _mm256_stream_si256((__m256i*)((char*)dst + o), _mm256_load_si256((const __m256i*)((const char*)src + o)));
  if (body < bytes) {
    memcpy((char*)dst + body, (const char*)src + body, bytes - body);
    writeback_range((char*)dst + body, bytes - body);
  }
}
#endif

template <typename D>
int pack(const void* const* cols, const int32_t* dtype, const int64_t* width, const int64_t* ld, int32_t n_cols, int64_t row_start,
         int64_t rows, D* dst, int64_t dst_ld, int32_t n_threads) {
  if (rows == 0) return HRB_OK;
  std::lock_guard<std::mutex> job(g_pack_mu);
  Pool& p = pool();
  int nt = n_threads <= 0 ? p.size() : std::min(n_threads, p.size());
  if (rows < 4096) nt = 1;
  const int64_t per = ((rows + nt - 1) / nt + 511) / 512 * 512;  // thread ranges start on 512-row block boundaries (aligned streaming stores)
  // row-major walk: a destination row (a few hundred bytes) is written in one go while every source column is read as its own
  // sequential stream -- the per-column walk this replaces spent ~6 ns per element on strided stores
  bool simple = true;  // every column is one element wide and already has the destination's element size (int32->int32, f32->f32)
  bool unit_stride = true;
  for (int c = 0; c < n_cols; ++c) {
    simple = simple && width[c] == 1 && dtype[c] == (std::is_same<D, int32_t>::value ? 0 : 2);
    unit_stride = unit_stride && ld[c] == 1;
  }
  auto work = [&](int t) {
    const int64_t a = row_start + t * per, b = std::min(row_start + rows, a + per);
    if (a >= b) return;
#if defined(__x86_64__)
    if (simple && unit_stride && n_cols >= 8 && sizeof(D) == 4 && have_avx2()) {
      constexpr int64_t RB = 512;  // rows per block: 512 x 39 x 4 B = 78 KB stays in L2 while it is assembled
      const uint32_t* const* c32 = reinterpret_cast<const uint32_t* const*>(cols);
      uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
      // Dense destination rows + 64-byte aligned blocks: assemble the block in a per-thread scratch and stream it out with
      // non-temporal stores -- the destination never enters the caches (no read-for-ownership, nothing left dirty for the DMA).
      const bool stream = dst_ld == n_cols && ((uintptr_t)dst & 63) == 0 && ((RB * n_cols * 4) & 63) == 0 && ((a - row_start) % RB) == 0 &&
                          g_writeback.load(std::memory_order_relaxed);
      uint32_t* scratch = stream ? thread_scratch((size_t)RB * n_cols) : nullptr;
      for (int64_t r0 = a; r0 < b; r0 += RB) {
        const int64_t r1 = std::min(b, r0 + RB);
        uint32_t* out = d32 + (r0 - row_start) * dst_ld;
        if (scratch) {
          pack_rows_avx2(c32, n_cols, r0, r1, r0, scratch, n_cols);
          stream_copy(out, scratch, (size_t)(r1 - r0) * n_cols * 4);
        } else {
          pack_rows_avx2(c32, n_cols, r0, r1, row_start, d32, dst_ld);
          writeback_range(out, (size_t)((r1 - r0 - 1) * dst_ld + n_cols) * sizeof(D));
        }
      }
      asm volatile("sfence" ::: "memory");
      return;
    }
#endif
    if (simple) {
      // blocks of 256 rows: every source column is read as one sequential 1 KB run (prefetcher-friendly; 39 interleaved streams
      // are not), the 256 x n_cols destination block stays in L1/L2 while its columns are filled
      constexpr int64_t RB = 256;
      for (int64_t r0 = a; r0 < b; r0 += RB) {
        const int64_t r1 = std::min(b, r0 + RB);
        D* d0 = dst + (r0 - row_start) * dst_ld;
        for (int c = 0; c < n_cols; ++c) {
          const D* s = static_cast<const D*>(cols[c]);
          const int64_t l = ld[c];
          D* d = d0 + c;
          for (int64_t r = r0; r < r1; ++r, d += dst_ld) *d = s[r * l];
        }
        writeback_range(d0, (size_t)((r1 - r0 - 1) * dst_ld + n_cols) * sizeof(D));
      }
#if defined(__x86_64__)
      asm volatile("sfence" ::: "memory");
#endif
      return;
    }
    for (int64_t r = a; r < b; ++r) {
      D* d = dst + (r - row_start) * dst_ld;
      for (int c = 0; c < n_cols; ++c) {
        const int64_t w = width[c];
        switch (dtype[c]) {
          case 0: { const int32_t* s = static_cast<const int32_t*>(cols[c]) + r * ld[c]; for (int64_t k = 0; k < w; ++k) d[k] = (D)s[k]; break; }
          case 1: { const int64_t* s = static_cast<const int64_t*>(cols[c]) + r * ld[c]; for (int64_t k = 0; k < w; ++k) d[k] = (D)s[k]; break; }
          case 2: { const float* s = static_cast<const float*>(cols[c]) + r * ld[c]; for (int64_t k = 0; k < w; ++k) d[k] = (D)s[k]; break; }
          default: { const double* s = static_cast<const double*>(cols[c]) + r * ld[c]; for (int64_t k = 0; k < w; ++k) d[k] = (D)s[k]; break; }
        }
        d += w;
      }
    }
    writeback_range(dst + (a - row_start) * dst_ld, (size_t)((b - a) * dst_ld) * sizeof(D));
#if defined(__x86_64__)
    asm volatile("sfence" ::: "memory");
#endif
  };
  if (nt == 1) work(0); else p.run(nt, work);
  return HRB_OK;
}

int check(const char* who, const void* const* cols, const int32_t* dtype, const int64_t* width, const int64_t* ld, int32_t n_cols,
          int64_t row_start, int64_t rows, const void* dst, int64_t dst_ld) {
  HRB_REQUIRE(n_cols >= 0 && rows >= 0 && row_start >= 0 && (rows == 0 || dst != nullptr), "%s: bad argument", who);
  int64_t total = 0;
  for (int c = 0; c < n_cols; ++c) {
    HRB_REQUIRE(cols && dtype && width && ld && cols[c] != nullptr && width[c] > 0 && ld[c] >= width[c] && dtype[c] >= 0 && dtype[c] <= 3,
                "%s: column %d is malformed", who, c);
    total += width[c];
  }
  HRB_REQUIRE(dst_ld >= total, "%s: dst_ld %lld < %lld packed columns", who, (long long)dst_ld, (long long)total);
  return HRB_OK;
}

}  // namespace

HRB_API int hrb_host_pack_i32(const void* const* cols_host, const int32_t* dtype_host, const int64_t* width_host, const int64_t* ld_host,
                              int32_t n_cols, int64_t row_start, int64_t rows, int32_t* dst_host, int64_t dst_ld, int32_t n_threads) {
  int rc = check("hrb_host_pack_i32", cols_host, dtype_host, width_host, ld_host, n_cols, row_start, rows, dst_host, dst_ld);
  if (rc != HRB_OK) return rc;
  return pack<int32_t>(cols_host, dtype_host, width_host, ld_host, n_cols, row_start, rows, dst_host, dst_ld, n_threads);
}

HRB_API int hrb_host_pack_f32(const void* const* cols_host, const int32_t* dtype_host, const int64_t* width_host, const int64_t* ld_host,
                              int32_t n_cols, int64_t row_start, int64_t rows, float* dst_host, int64_t dst_ld, int32_t n_threads) {
  int rc = check("hrb_host_pack_f32", cols_host, dtype_host, width_host, ld_host, n_cols, row_start, rows, dst_host, dst_ld);
  if (rc != HRB_OK) return rc;
  return pack<float>(cols_host, dtype_host, width_host, ld_host, n_cols, row_start, rows, dst_host, dst_ld, n_threads);
}

HRB_API int hrb_host_pack_set_writeback(int32_t on) {
  g_writeback.store(on ? 1 : 0);
  return cache_writeback_kind();
}
