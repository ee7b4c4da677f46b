// gemm_tc.cu -- tcgen05 (5th-gen tensor core) GEMM with 3xTF32 error compensation, used by the
// Dense layers of the DNN towers (layers/core.py:53-78) when the problem is a real GEMM (M >= 1024).
//
//   C[M,N] = A[M,K] * Bt[N,K]^T      A, Bt row-major with the reduction dim contiguous ("K-major")
//
// fp32 parity (<= 1e-5 forward, <= 1e-4 gradients) rules out plain TF32 (10-bit mantissa).  Every fp32
// operand tile is split into hi = x & 0xFFFFE000 (what kind::tf32 reads of a raw fp32 word -- the low 13 mantissa
// bits are ignored, verified against explicit masking) and lo = x - hi (written to a second smem tile), and three
// MMAs accumulate lo*hi + hi*lo + hi*hi in TMEM (fp32): error ~2^-21, like FFMA.
//
// Persistent, warp-specialised (512 threads, 1 CTA/SM):
//   warp 0      TMA producer   cp.async.bulk.tensor.2d (SWIZZLE_128B boxes of 32 fp32 = 128 B x rows)
//   warp 1      MMA issuer     tcgen05.mma.cta_group::1.kind::tf32, 3 x (BK/8) instructions per k-block
//   warp 2      TMEM allocator (2 accumulator stages x BN columns)
//   warps 4-7   epilogue       tcgen05.ld 32x32b.x32 -> bias/activation/activation-gradient -> swizzled smem
//                              staging -> TMA stores (cp.async.bulk.tensor, full 128-byte lines, OOB clipped):
//                              row-major C and, optionally, the transposed copy Ct for the next bwd_w
//   warps 8-15  converters     hi/lo split of each landed stage, in place, fence.proxy.async
// Pipelines: full (TMA->conv), conv (conv->MMA), empty (MMA->TMA, tcgen05.commit), tmem_full / tmem_empty.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace hrb {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;  // 32 fp32 = 128 bytes = one SWIZZLE_128B span
constexpr int STAGES_SMEM_A = 3;  // A hi/lo in shared memory: 64 KB per stage
constexpr int STAGES_TMEM_A = 4;  // A hi/lo in tensor memory: 48 KB per stage, 64 TMEM columns per stage
constexpr int THREADS = 512;
constexpr int EPI_WARP0 = 4, CONV_WARP0 = 8;
constexpr int CONV_THREADS = 256;
constexpr uint32_t STAGE_C_BYTES = 128 * 32 * 4;  // one 128 x 32 fp32 output sub-tile

enum { EPI_BIAS_ACT = 0, EPI_ACT_GRAD = 1, EPI_PLAIN = 2 };

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(s32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug becomes a trap (CUDA error) instead of a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity))
    if (clock64() - t0 > 8000000000LL) __trap();
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(dst)),
      "l"(map), "r"(s32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(s32(src)), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(s32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A operand read from tensor memory (lane = row of the 128-row tile, one 32-bit column per tf32 element)
__device__ __forceinline__ void umma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
      "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// ---- CTA pair (cta_group::2) ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(s32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint64_t* bar, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(s32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ bool mbar_try_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(s32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_cluster(bar, parity))
    if (clock64() - t0 > 8000000000LL) __trap();
}
// M = 256 over the pair: each CTA contributes its 128 rows of A (tensor memory) and its half of B (shared memory)
__device__ __forceinline__ void umma2_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the arrival lands on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major) | [32,46) SBO>>4 = 1024 B between
//   8-row groups | [46,48) version = 1 (sm100) | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(const void* smem) {
  uint64_t d = 0;
  d |= (uint64_t)((s32(smem) >> 4) & 0x3FFF);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct Args {
  float* C;
  float* Ct;           // optional transposed copy [N][ldct]
  const float* bias;   // EPI_BIAS_ACT
  const float* aprev;  // EPI_ACT_GRAD: previous layer's post-activation output [M][ldap]
  int64_t ldc, ldct, ldap;
  int64_t M;
  int32_t N, K;
  int32_t act;
  int32_t splits;      // split over K (EPI_PLAIN): slice z writes C + z*M*ldc
  int32_t kb_per_split;
  uint32_t* mask_out;  // EPI_BIAS_ACT + relu: bit j of word [m][n0/32] = (y[m][n0+j] > 0), for the backward
  const uint32_t* mask_in;  // EPI_ACT_GRAD + relu: the same words instead of re-reading the activations
  int64_t mask_ld;     // words per row
  float* bsum;         // EPI_PLAIN: bsum[z*N + n] = sum over this split's k of Bt[n][k] (the bias gradient when Bt = dz^T), or NULL
  long long* trace;    // HRB_TC_TRACE: clock64 stamps of the first 64 k-blocks of CTAs 0/1 (perf debugging)
  int32_t debug;       // HRB_TC_DEBUG bit mask (perf experiments only): 1 no stores, 2 no conversion, 4 no MMA, 8 no TMA, 16 also write hi back
  // EPI_BIAS_ACT with a bias per GROUP of rows (row m takes bias[(m / bias_group) * bias_ld + n]): the per-sample query term of the
  // local-activation MLP's first layer (hrb_dense_fwd_t_grouped)
  int32_t bias_group;
  int64_t bias_ld;
};

// ATM: the A tile's hi/lo halves live in TENSOR MEMORY (written by the converter warps with tcgen05.st) instead of shared
// memory.  The kernel is bound by shared-memory bandwidth (TMA writes + converter read/write + 6 operand-tile reads per
// k-block by the three MMAs, ~245 B/clk wanted against 128 B/clk): with A in TMEM the MMAs read only B from shared memory
// and the converter no longer writes A_lo there (-36 % shared-memory traffic per k-block).
//
// PAIR (implies ATM): two CTAs of a cluster (one TPC) run ONE 256 x BN tile with tcgen05.mma.cta_group::2.  Each CTA loads and
// converts its own 128 rows of A (-> its tensor memory) and HALF of the B tile (BN/2 rows -> its shared memory); the leader CTA's
// MMA thread issues for both, tcgen05.commit multicasts the "stage free" / "accumulator full" arrivals to both CTAs, the peer's
// converter and epilogue warps arrive on the leader's barriers through the cluster address space.  Per CTA and k-block this moves
// 24 KB instead of 32 KB from L2 and 80 KB instead of 128 KB through shared memory.
// Perf-debugging hooks (ablation bits, clock64 stamps, alternative variants selected by HRB_TC_* environment variables) exist
// only in a -DHRB_DEVTOOLS build; the shipped library has none of them.
#ifdef HRB_DEVTOOLS
#define HRB_DBG(b) (g.debug & (b))
#define HRB_TRACE(role, it_) \
  if (g.trace != nullptr && blockIdx.x < 2 && (it_) < 64) g.trace[((blockIdx.x * 8 + (role)) << 6) + (it_)] = clock64();
#else
#define HRB_DBG(b) 0
#define HRB_TRACE(role, it_)
#endif

// AMN (split-K weight gradient only): the A operand is given UNtransposed -- x[samples][features] as the forward stored it.  The
// TMA box is 32 samples x 128 features (no swizzle, 512-byte rows) and the converter thread of feature row r picks x[k][r] for its 16
// samples (consecutive lanes = consecutive words: conflict-free), so no x^T copy has to exist in HBM.
template <int BN, int EPI, bool ATM, bool PAIR, bool AMN>
__global__ void __launch_bounds__(THREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                             const __grid_constant__ CUtensorMap map_b,
                                                             const __grid_constant__ CUtensorMap map_c,
                                                             const __grid_constant__ CUtensorMap map_ct, Args g) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  static_assert(!PAIR || ATM, "the CTA-pair kernel keeps A in tensor memory");
  constexpr int STAGES = ATM ? STAGES_TMEM_A : STAGES_SMEM_A;
  constexpr int BNL = PAIR ? BN / 2 : BN;  // rows of B held by this CTA
  constexpr uint32_t A_BYTES = BM * BK * 4, B_BYTES = BNL * BK * 4;
  constexpr uint32_t B_OFF = ATM ? A_BYTES : 2 * A_BYTES;
  constexpr uint32_t STAGE_BYTES = B_OFF + 2 * B_BYTES;  // A (raw = hi) | [A_lo] | B (raw = hi) | B_lo
  constexpr uint32_t ACC_COLS = 2 * BN;                  // two accumulator stages
  constexpr uint32_t A_COLS = 2 * BK;                    // ATM: per stage, hi columns then lo columns
  constexpr uint32_t TMEM_NEED = ACC_COLS + (ATM ? STAGES * A_COLS : 0);
  constexpr uint32_t TMEM_COLS = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
  static_assert(TMEM_NEED <= 512, "tensor memory budget");
  // 1024-byte alignment of every tile (SWIZZLE_128B atoms are 8 rows x 128 B)
  // (pointer arithmetic on the __shared__ array keeps the address space known: LDS/STS, not generic LD/ST)
  unsigned char* tiles = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t full_bar[STAGES], conv_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler: role branches do not diverge
  const int lane = threadIdx.x & 31;
  const uint32_t pair_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = pair_rank == 0;
  constexpr int BMT = PAIR ? 2 * BM : BM;                    // rows of one work item
  const int64_t w_first = PAIR ? (blockIdx.x >> 1) : blockIdx.x, w_step = PAIR ? (gridDim.x >> 1) : gridDim.x;
  const int m_tiles = (int)((g.M + BMT - 1) / BMT), n_tiles = (g.N + BN - 1) / BN;
  const int64_t n_work = (int64_t)m_tiles * n_tiles * g.splits;
  const int kb_total = (g.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&conv_bar[s], PAIR ? CONV_THREADS / 32 + 1 : CONV_THREADS / 32);  // one arrival per converter warp (+ one for the peer CTA)
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], PAIR ? 8 : 4);                            // one arrival per epilogue warp (both CTAs)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // one warp allocates 2 accumulator stages (+ the A stages), a power of two >= 32 columns
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_smem)), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tmem_base_smem)), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // both CTAs' barriers are initialised before anything arrives on them remotely
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_smem, 0);

  // work item -> (m tile, n tile, k split); n fastest so neighbouring CTAs share the A rows in L2.  The last n tile is NARROW
  // (the MMA runs with N = its real width rounded to 16, e.g. 48 of 128 columns at N = 429), so tiles differ in cost: the n order
  // inside a group is rotated by the pass number so that every CTA (stride w_step) meets all n tiles in turn.
  auto decode = [&](int64_t w, int& mt, int& nt, int& z) {
    z = (int)(w / ((int64_t)m_tiles * n_tiles));
    const int64_t r = w - (int64_t)z * m_tiles * n_tiles;
    mt = (int)(r / n_tiles);
    nt = (int)(r - (int64_t)mt * n_tiles);
    nt = (int)((nt + ((int64_t)mt * n_tiles) / w_step) % n_tiles);
  };
  // columns of n tile nt that exist, rounded up to the MMA's N granularity (16); this CTA's share of the B rows
  auto n_width = [&](int nt) { return min(BN, (g.N - nt * BN + 15) & ~15); };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t w = w_first; w < n_work; w += w_step) {
        int mt, nt, z;
        decode(w, mt, nt, z);
        const int kb0 = z * g.kb_per_split, kb1 = min(kb_total, kb0 + g.kb_per_split);
        const int b_half = PAIR ? n_width(nt) / 2 : 0;  // the pair splits the tile's REAL width: rank 1 starts at column n_width/2
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
          unsigned char* st = tiles + (size_t)s * STAGE_BYTES;
          if (HRB_DBG(8)) {
            mbar_arrive(&full_bar[s]);
            continue;
          }
          HRB_TRACE(0, it)
          mbar_expect_tx(&full_bar[s], A_BYTES + B_BYTES);
          if (AMN)
            tma_load_2d(st, &map_a, &full_bar[s], mt * BMT + (int)pair_rank * BM, kb * BK);  // (feature, sample) coordinates
          else
            tma_load_2d(st, &map_a, &full_bar[s], kb * BK, mt * BMT + (int)pair_rank * BM);
          tma_load_2d(st + B_OFF, &map_b, &full_bar[s], kb * BK, nt * BN + (int)pair_rank * b_half);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop converged and every operand is derived from warp-uniform values, so the descriptors sit in
    // uniform registers; ONE elected lane issues.  (Under `if (lane == 0)` the compiler wrapped every tcgen05.mma in a
    // R2UR/ELECT waterfall of ~80 cycles -- longer than the 64 cycles the instruction occupies the tensor pipe.)
    if (leader) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a=b=TF32 [7,10)/[10,13),
      // K-major A and B (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc0 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BMT >> 4) << 24);
      uint32_t it = 0, tcount = 0;
      for (int64_t w = w_first; w < n_work; w += w_step, ++tcount) {
        int mt, nt, z;
        decode(w, mt, nt, z);
        const uint32_t idesc = idesc0 | ((uint32_t)(n_width(nt) >> 3) << 17);
        const int kb0 = z * g.kb_per_split, kb1 = min(kb_total, kb0 + g.kb_per_split);
        const int a = tcount & 1;
        if (PAIR)
          mbar_wait_cluster(&tempty_bar[a], ((tcount >> 1) & 1) ^ 1);
        else
          mbar_wait(&tempty_bar[a], ((tcount >> 1) & 1) ^ 1);  // epilogue drained this accumulator
        tc_fence_after();
        if (lane == 0) { HRB_TRACE(7, tcount) }
        const uint32_t tmem_d = tmem_base + (uint32_t)(a * BN);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          if (PAIR)
            mbar_wait_cluster(&conv_bar[s], (it / STAGES) & 1);
          else
            mbar_wait(&conv_bar[s], (it / STAGES) & 1);
          tc_fence_after();
          unsigned char* st = tiles + (size_t)s * STAGE_BYTES;
          const uint64_t b_hi = make_desc(st + B_OFF), b_lo = make_desc(st + B_OFF + B_BYTES);
          const uint32_t first = kb > kb0 ? 1u : 0u;
          if (elect_one()) {
            HRB_TRACE(3, it)
            if (ATM) {
              const uint32_t ta_hi = tmem_base + ACC_COLS + (uint32_t)s * A_COLS, ta_lo = ta_hi + BK;
              if (!(HRB_DBG(4)))
#pragma unroll
              for (int kk = 0; kk < BK / 8; ++kk) {  // UMMA_K = 8 tf32: 8 TMEM columns of A, 32 bytes of B
                const uint64_t o = (uint64_t)(kk * 2);
                if (PAIR) {
                  umma2_tf32_ta(tmem_d, ta_lo + kk * 8, b_hi + o, idesc, kk > 0 ? 1u : first);
                  umma2_tf32_ta(tmem_d, ta_hi + kk * 8, b_lo + o, idesc, 1u);
                  umma2_tf32_ta(tmem_d, ta_hi + kk * 8, b_hi + o, idesc, 1u);
                } else {
                  umma_tf32_ta(tmem_d, ta_lo + kk * 8, b_hi + o, idesc, kk > 0 ? 1u : first);
                  umma_tf32_ta(tmem_d, ta_hi + kk * 8, b_lo + o, idesc, 1u);
                  umma_tf32_ta(tmem_d, ta_hi + kk * 8, b_hi + o, idesc, 1u);
                }
              }
            } else {
              const uint64_t a_hi = make_desc(st), a_lo = make_desc(st + A_BYTES);
              if (!(HRB_DBG(4)))
#pragma unroll
              for (int kk = 0; kk < BK / 8; ++kk) {  // UMMA_K = 8 tf32 = 32 bytes: +2 in the (>>4) start-address field
                const uint64_t o = (uint64_t)(kk * 2);
                umma_tf32(tmem_d, a_lo + o, b_hi + o, idesc, kk > 0 ? 1u : first);
                umma_tf32(tmem_d, a_hi + o, b_lo + o, idesc, 1u);
                umma_tf32(tmem_d, a_hi + o, b_hi + o, idesc, 1u);
              }
            }
            if (PAIR) umma2_commit(&empty_bar[s]); else umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs retire
            if (kb + 1 == kb1) {  // accumulator complete -> epilogue (same thread as the MMAs: commit tracks the issuing thread)
              if (PAIR) umma2_commit(&tfull_bar[a]); else umma_commit(&tfull_bar[a]);
            }
            HRB_TRACE(4, it)
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= CONV_WARP0) {
    // ===================== converters: hi/lo split in place =====================
    const int t = threadIdx.x - CONV_WARP0 * 32;  // 0..255
    constexpr int B_PER_THREAD = (int)(B_BYTES / 16) / CONV_THREADS;  // float4s of the B tile per thread: rows t/8 + 32 j
    const bool want_bsum = EPI == EPI_PLAIN && g.bsum != nullptr;
    uint32_t it = 0;
    for (int64_t w = w_first; w < n_work; w += w_step) {
      int mt, nt, z;
      decode(w, mt, nt, z);
      const int kb0 = z * g.kb_per_split, kb1 = min(kb_total, kb0 + g.kb_per_split);
      float bs[B_PER_THREAD];
#pragma unroll
      for (int j = 0; j < B_PER_THREAD; ++j) bs[j] = 0.f;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&full_bar[s], (it / STAGES) & 1);
        if (t == 0) { HRB_TRACE(1, it) }
        unsigned char* st = tiles + (size_t)s * STAGE_BYTES;
        if (ATM) {
          // A: this thread owns row (warp%4)*32+lane of the tile (the TMEM lanes its warp may touch) and 16 of the 32 k-columns;
          // row r of a SWIZZLE_128B tile keeps its 16-byte chunk c at chunk c ^ (r % 8)
          if (!(HRB_DBG(2))) {
            const int qrow = warp & 3, half = (warp - CONV_WARP0) >> 2;
            const int row = qrow * 32 + lane;
            const float4* src = reinterpret_cast<const float4*>(st + row * 128);
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              float4 x;
              if (AMN) {  // tile is [32 samples][128 features] plain: element (row, k) sits at k * 512 + row * 4 bytes
                const float* col = reinterpret_cast<const float*>(st) + row;
                x = make_float4(col[(half * 16 + c * 4 + 0) * 128], col[(half * 16 + c * 4 + 1) * 128], col[(half * 16 + c * 4 + 2) * 128],
                                col[(half * 16 + c * 4 + 3) * 128]);
              } else {
                x = src[(half * 4 + c) ^ (row & 7)];
              }
              const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t h = __float_as_uint(xs[j]) & 0xFFFFE000u;
                hi[c * 4 + j] = h;
                lo[c * 4 + j] = __float_as_uint(xs[j] - __uint_as_float(h));
              }
            }
            const uint32_t ta = tmem_base + ((uint32_t)(qrow * 32) << 16) + ACC_COLS + (uint32_t)s * A_COLS + (uint32_t)(half * 16);
            tmem_st16(ta, hi);
            tmem_st16(ta + BK, lo);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          }
        } else {
        // element-wise, so the swizzled placement is irrelevant: lo lands at the same offset as hi
        if (!(HRB_DBG(2)))
#pragma unroll 4
        for (int i = t; i < (int)(A_BYTES / 16); i += CONV_THREADS) {
          float4* p = reinterpret_cast<float4*>(st) + i;
          float4 x = *p, h;
          h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
          h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
          h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
          h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
          if (HRB_DBG(16)) *p = h;  // the MMA reads only the TF32 bits of the raw fp32 operand: writing hi back is redundant
          reinterpret_cast<float4*>(st + A_BYTES)[i] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
        }
        }
        if (!(HRB_DBG(2)))
#pragma unroll
        for (int j = 0; j < B_PER_THREAD; ++j) {
          const int i = t + j * CONV_THREADS;
          float4* p = reinterpret_cast<float4*>(st + B_OFF) + i;
          float4 x = *p, h;
          h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
          h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
          h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
          h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
          if (HRB_DBG(16)) *p = h;  // the MMA reads only the TF32 bits of the raw fp32 operand: writing hi back is redundant
          reinterpret_cast<float4*>(st + B_OFF + B_BYTES)[i] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
          if (EPI == EPI_PLAIN) bs[j] += (x.x + x.y) + (x.z + x.w);  // row sums of Bt ride along (bias gradient)
        }
        if (ATM) tc_fence_before();  // the tcgen05.st above are ordered before the MMA issuer's tcgen05.mma by the barrier
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to UMMA
        if (t == 0) { HRB_TRACE(2, it) }
        if (PAIR && !leader) {
          asm volatile("bar.sync 2, 256;" ::: "memory");  // this CTA's 256 converter threads are done with the stage
          if (t == 0) {
            HRB_TRACE(3, it)
            // relaxed: the writes were fenced to the async proxy by every thread and ordered by the bar.sync above; a
            // release at cluster scope stalls this thread for 1000-2000 cycles per k-block (measured) and serialises the peer
            if (HRB_DBG(64)) mbar_arrive_remote(&conv_bar[s], 0); else mbar_arrive_remote_relaxed(&conv_bar[s], 0);
            HRB_TRACE(4, it)
          }
        } else {
          __syncwarp();  // orders the other lanes' fenced writes before lane 0's releasing arrive: 8 arrivals per stage instead of 256
          if (lane == 0) mbar_arrive(&conv_bar[s]);
        }
      }
      if (EPI == EPI_PLAIN) {  // 8 consecutive lanes hold the 8 chunks of one 128-byte row: fixed-order tree, deterministic
#pragma unroll
        for (int j = 0; j < B_PER_THREAD; ++j) {
          float v = bs[j];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          const int b_half = PAIR ? n_width(nt) / 2 : BN, row = (t >> 3) + 32 * j;
          const int n = nt * BN + (int)pair_rank * b_half + row;
          if (want_bsum && mt == 0 && (t & 7) == 0 && row < b_half && n < g.N) g.bsum[(size_t)z * g.N + n] = v;
        }
      }
    }
  } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + 4) {
    // ===================== epilogue =====================
    const int q = warp - EPI_WARP0;  // == warp % 4: the TMEM lane quadrant this warp may read
    const int et = q * 32 + lane;    // row of the 128-row tile owned by this thread
    // staging: [128][32] fp32 with SWIZZLE_128B rows (C) + [32][128] plain (Ct); the pair kernel has the shared memory for two sets,
    // so a sub-tile is staged while the TMA store of the previous one still reads the other set
    constexpr int EPI_BUFS = PAIR ? 2 : 1;
    float* sC0 = reinterpret_cast<float*>(tiles + (size_t)STAGES * STAGE_BYTES);
    uint32_t tcount = 0, sub = 0;
    for (int64_t w = w_first; w < n_work; w += w_step, ++tcount) {
      int mt, nt, z;
      decode(w, mt, nt, z);
      const int a = tcount & 1;
      mbar_wait(&tfull_bar[a], (tcount >> 1) & 1);
      tc_fence_after();
      if (et == 0) { HRB_TRACE(5, tcount) }
      const int64_t m = (int64_t)mt * BMT + (int64_t)pair_rank * BM + et;
      const int m_row0 = mt * BMT + (int)pair_rank * BM;
      const bool row_ok = m < g.M;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
            "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
              "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
              "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int n0 = nt * BN + c0;
        if (n0 >= g.N || (HRB_DBG(1))) continue;  // uniform: the whole sub-tile is outside the matrix
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (EPI == EPI_BIAS_ACT) {
          if (g.bias != nullptr && g.bias_group > 0) {
            if (row_ok) {
              const float* gb = g.bias + (m / g.bias_group) * g.bias_ld + n0;
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += (n0 + j < g.N) ? __ldg(gb + j) : 0.f;
            }
          } else if (g.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += (n0 + j < g.N) ? __ldg(g.bias + n0 + j) : 0.f;
          }
          if (g.act == HRB_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            if (g.mask_out != nullptr && row_ok) {
              uint32_t word = 0;
#pragma unroll
              for (int j = 0; j < 32; ++j) word |= (v[j] > 0.f ? 1u : 0u) << j;
              g.mask_out[m * g.mask_ld + (n0 >> 5)] = word;
            }
          } else if (g.act != HRB_ACT_LINEAR) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = act_apply(g.act, v[j]);
          }
        } else if (EPI == EPI_ACT_GRAD) {
          if (g.mask_in != nullptr && g.act == HRB_ACT_RELU) {
            const uint32_t word = row_ok ? __ldg(g.mask_in + m * g.mask_ld + (n0 >> 5)) : 0u;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = ((word >> j) & 1u) ? v[j] : 0.f;
          } else if (g.aprev != nullptr && row_ok) {
            const float* ap = g.aprev + m * g.ldap + n0;
#pragma unroll
            for (int j4 = 0; j4 < 32; j4 += 4) {
              float4 y = make_float4(1.f, 1.f, 1.f, 1.f);
              if (n0 + j4 + 3 < g.N) {
                y = __ldg(reinterpret_cast<const float4*>(ap + j4));
              } else {
                if (n0 + j4 < g.N) y.x = __ldg(ap + j4);
                if (n0 + j4 + 1 < g.N) y.y = __ldg(ap + j4 + 1);
                if (n0 + j4 + 2 < g.N) y.z = __ldg(ap + j4 + 2);
              }
              if (g.act == HRB_ACT_RELU) {
                v[j4] = y.x > 0.f ? v[j4] : 0.f; v[j4 + 1] = y.y > 0.f ? v[j4 + 1] : 0.f;
                v[j4 + 2] = y.z > 0.f ? v[j4 + 2] : 0.f; v[j4 + 3] = y.w > 0.f ? v[j4 + 3] : 0.f;
              } else {
                v[j4] *= act_grad_from_out(g.act, y.x); v[j4 + 1] *= act_grad_from_out(g.act, y.y);
                v[j4 + 2] *= act_grad_from_out(g.act, y.z); v[j4 + 3] *= act_grad_from_out(g.act, y.w);
              }
            }
          }
        }
        // a staging set is free once the TMA stores issued from it (EPI_BUFS sub-tiles ago) have READ it
        float* sC = sC0 + (size_t)(sub % EPI_BUFS) * (2 * STAGE_C_BYTES / 4);
        float* sCt = sC + STAGE_C_BYTES / 4;
        ++sub;
        if (et == 0) {
            if (EPI_BUFS == 2)
              asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        epi_bar();
#pragma unroll
        for (int c = 0; c < 8; ++c)  // row `et`, 16-byte chunk c lands at chunk (c ^ (et % 8)): SWIZZLE_128B
          *reinterpret_cast<float4*>(sC + et * 32 + ((c ^ (et & 7)) << 2)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        if (g.Ct != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) sCt[j * 128 + et] = v[j];
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        epi_bar();
        if (et == 0) {
          tma_store_3d(&map_c, sC, n0, m_row0, z);
          if (g.Ct != nullptr) tma_store_2d(&map_ct, sCt, m_row0, n0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      tc_fence_before();
      if (et == 0) { HRB_TRACE(6, tcount) }
      __syncwarp();
      if (lane == 0) {
        if (PAIR && !leader)
          mbar_arrive_remote(&tempty_bar[a], 0);
        else
          mbar_arrive(&tempty_bar[a]);
      }
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // nobody leaves (or frees tensor memory) while the peer may still arrive here
  if (warp == 2) {
    if (PAIR)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    else
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp32 row-major [rows][cols] with leading dim ld; box = box_cols x box_rows
static int make_map2(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
                     CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail(HRB_CUDA_ERROR, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(HRB_CUDA_ERROR, "cuTensorMapEncodeTiled(2d) failed with CUresult %d", (int)r);
  return HRB_OK;
}
// 3-D [slices][rows][cols] output map (slices = split-K partials), box = 32 x 128 x 1, SWIZZLE_128B
static int make_map3(CUtensorMap* map, const float* base, int64_t slices, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail(HRB_CUDA_ERROR, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)slices};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 4, (cuuint64_t)rows * (cuuint64_t)ld * 4};
  cuuint32_t box[3] = {32, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(HRB_CUDA_ERROR, "cuTensorMapEncodeTiled(3d) failed with CUresult %d", (int)r);
  return HRB_OK;
}

static void dump_trace(long long* d_trace, cudaStream_t st) {
  static int dumped = 0;
  if (d_trace == nullptr || dumped >= 2) return;
  ++dumped;
  static long long h[2 * 8 * 64];
  cudaStreamSynchronize(st);
  cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[8] = {"tma_issue", "conv_full", "conv_done", "mma_ready", "mma_commit", "epi_start", "epi_end", "mma_tile"};
  for (int c = 0; c < 2; ++c) {
    const long long t0 = h[(c * 8 + 0) << 6];
    for (int r = 0; r < 8; ++r) {
      fprintf(stderr, "cta%d %-10s", c, names[r]);
      for (int i = 0; i < 40; ++i) fprintf(stderr, " %lld", h[((c * 8 + r) << 6) + i] ? h[((c * 8 + r) << 6) + i] - t0 : -1);
      fprintf(stderr, "\n");
    }
  }
}

// The cta_group::2 kernel (two CTAs of a cluster share one 256 x 128 tile) is the default; HRB_TC_PAIR=0 selects the single-CTA one
static bool pair_enabled() {
#ifdef HRB_DEVTOOLS
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HRB_TC_PAIR");
    v = (e != nullptr && atoi(e) == 0) ? 0 : 1;
    const char* a = getenv("HRB_TC_A");
    if (a != nullptr && strcmp(a, "smem") == 0) v = 0;
  }
  return v != 0;
#else
  return true;
#endif
}

template <int BN, int EPI, bool AMN = false>
static int launch(const float* A, int64_t lda, const float* Bt, int64_t ldb, const Args& g_in, cudaStream_t st) {
  Args g = g_in;
  long long* d_trace = nullptr;
#ifdef HRB_DEVTOOLS
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("HRB_TC_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  g.debug = dbg;
  static long long* d_trace_buf = nullptr;
  static int want_trace = -1;
  if (want_trace < 0) {
    want_trace = getenv("HRB_TC_TRACE") != nullptr ? 1 : 0;
    if (want_trace) cudaMalloc((void**)&d_trace_buf, 2 * 8 * 64 * sizeof(long long));
  }
  d_trace = d_trace_buf;
  if (d_trace != nullptr) cudaMemsetAsync(d_trace, 0, 2 * 8 * 64 * sizeof(long long), st);
  // HRB_TC_A=smem keeps the A tile's hi/lo halves in shared memory (the first version of the kernel; not for the untransposed-A form)
  static int a_tmem = -1;
  if (a_tmem < 0) {
    const char* e = getenv("HRB_TC_A");
    a_tmem = (e != nullptr && strcmp(e, "smem") == 0) ? 0 : 1;
  }
#else
  g.debug = 0;
  const int a_tmem = 1;
#endif
  g.trace = d_trace;
  if (AMN && !a_tmem) return fail(HRB_UNSUPPORTED, "tcgen05 GEMM: the untransposed-A weight gradient needs A in tensor memory (unset HRB_TC_A)");
  const bool pair = a_tmem && pair_enabled();
  CUtensorMap ma, mb, mc, mct;
  int rc = AMN ? make_map2(&ma, A, g.K, g.M, lda, BM, BK, CU_TENSOR_MAP_SWIZZLE_NONE)  // A = x[K samples][M features], box 128 x 32
               : make_map2(&ma, A, g.M, g.K, lda, BK, BM, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != HRB_OK) return rc;
  rc = make_map2(&mb, Bt, g.N, g.K, ldb, BK, pair ? BN / 2 : BN, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != HRB_OK) return rc;
  rc = make_map3(&mc, g.C, g.splits, g.M, g.N, g.ldc);
  if (rc != HRB_OK) return rc;
  if (g.Ct != nullptr) {
    rc = make_map2(&mct, g.Ct, g.N, g.M, g.ldct, 128, 32, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc != HRB_OK) return rc;
  } else {
    mct = mc;
  }
  constexpr size_t smem_s = (size_t)STAGES_SMEM_A * (2 * BM * BK * 4 + 2 * BN * BK * 4) + 2 * STAGE_C_BYTES + 1024;
  constexpr size_t smem_t = (size_t)STAGES_TMEM_A * (BM * BK * 4 + 2 * BN * BK * 4) + 2 * STAGE_C_BYTES + 1024;
  constexpr size_t smem_p = (size_t)STAGES_TMEM_A * (BM * BK * 4 + BN * BK * 4) + 4 * STAGE_C_BYTES + 1024;
  static_assert(smem_s <= 227 * 1024 && smem_t <= 227 * 1024, "shared memory budget");
  static bool attr_done = false;
  if (!attr_done) {
    HRB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI, false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
    HRB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI, true, false, AMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
    HRB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI, true, true, AMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));
    attr_done = true;
  }
  if (pair) {
    // one work item = a 256 x BN tile of a CTA pair (cluster of 2 = the two SMs of a TPC); persistent over sm_count()/2 pairs
    const int64_t work = (int64_t)((g.M + 2 * BM - 1) / (2 * BM)) * ((g.N + BN - 1) / BN) * g.splits;
    const int64_t pairs = work < sm_count() / 2 ? work : sm_count() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem_p;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    HRB_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, EPI, true, true, AMN>, ma, mb, mc, mct, g));
    HRB_LAUNCH_CHECK();
    dump_trace(d_trace, st);
    return HRB_OK;
  }
  const int64_t work = (int64_t)((g.M + BM - 1) / BM) * ((g.N + BN - 1) / BN) * g.splits;
  int64_t grid = work < sm_count() ? work : sm_count();
  if (a_tmem)
    gemm_tc_kernel<BN, EPI, true, false, AMN><<<(unsigned)grid, THREADS, smem_t, st>>>(ma, mb, mc, mct, g);
  else
    gemm_tc_kernel<BN, EPI, false, false, false><<<(unsigned)grid, THREADS, smem_s, st>>>(ma, mb, mc, mct, g);
  HRB_LAUNCH_CHECK();
  dump_trace(d_trace, st);
  return HRB_OK;
}

static bool tc_ok(const float* A, int64_t lda, const float* Bt, int64_t ldb, int64_t M, int32_t N, int32_t K) {
  // TMA needs 16-byte aligned bases and row pitches; tiny problems stay on the FFMA kernel
  return aligned16(A) && aligned16(Bt) && lda % 4 == 0 && ldb % 4 == 0 && M >= 512 && N >= 16 && K >= 16;
}

}  // namespace tc
}  // namespace hrb

using namespace hrb;

// Internal entry points (exported through dense.cu): all operands "TN" = reduction dim contiguous.
// C[M,N] = act(A[M,K] * Bt[N,K]^T + bias)  (+ transposed copy Ct[N,M])
int hrb_tc_gemm_bias_act(const float* a, int64_t lda, const float* bt, int64_t ldb, const float* bias, int64_t M, int32_t N, int32_t K,
                         int32_t act, float* c, int64_t ldc, float* ct, int64_t ldct, uint32_t* relu_mask, int64_t mask_ld, cudaStream_t st) {
  if (!tc::tc_ok(a, lda, bt, ldb, M, N, K) || !aligned16(c) || ldc % 4 != 0 || (ct != nullptr && (!aligned16(ct) || ldct % 4 != 0)))
    return fail(HRB_UNSUPPORTED, "tcgen05 GEMM: shape/alignment not covered");
  tc::Args g{c, ct, bias, nullptr, ldc, ldct, 0, M, N, K, act, 1, (K + tc::BK - 1) / tc::BK, relu_mask, nullptr, mask_ld, nullptr, nullptr, 0};
  return tc::launch<128, tc::EPI_BIAS_ACT>(a, lda, bt, ldb, g, st);
}
// C[M,N] = (A[M,K] * Bt[N,K]^T) * act'(aprev[M,N])  (+ transposed copy)
int hrb_tc_gemm_act_grad(const float* a, int64_t lda, const float* bt, int64_t ldb, int64_t M, int32_t N, int32_t K, const float* aprev,
                         int64_t ldap, int32_t act, float* c, int64_t ldc, float* ct, int64_t ldct, const uint32_t* relu_mask, int64_t mask_ld,
                         cudaStream_t st) {
  if (!tc::tc_ok(a, lda, bt, ldb, M, N, K) || !aligned16(c) || ldc % 4 != 0 || (aprev != nullptr && (!aligned16(aprev) || ldap % 4 != 0)) ||
      (ct != nullptr && (!aligned16(ct) || ldct % 4 != 0)))
    return fail(HRB_UNSUPPORTED, "tcgen05 GEMM: shape/alignment not covered");
  tc::Args g{c, ct, nullptr, aprev, ldc, ldct, ldap, M, N, K, act, 1, (K + tc::BK - 1) / tc::BK, nullptr, relu_mask, mask_ld, nullptr, nullptr, 0};
  return tc::launch<128, tc::EPI_ACT_GRAD>(a, lda, bt, ldb, g, st);
}
// C[M,N] = act(A[M,K] * Bt[N,K]^T + gbias[m / group, :])
int hrb_tc_gemm_groupbias_act(const float* a, int64_t lda, const float* bt, int64_t ldb, const float* gbias, int64_t ldgb, int32_t group,
                              int64_t M, int32_t N, int32_t K, int32_t act, float* c, int64_t ldc, cudaStream_t st) {
  if (!tc::tc_ok(a, lda, bt, ldb, M, N, K) || !aligned16(c) || ldc % 4 != 0) return fail(HRB_UNSUPPORTED, "tcgen05 GEMM: shape/alignment not covered");
  tc::Args g{c, nullptr, gbias, nullptr, ldc, 0, 0, M, N, K, act, 1, (K + tc::BK - 1) / tc::BK, nullptr, nullptr, 0, nullptr, nullptr, 0};
  g.bias_group = group;
  g.bias_ld = ldgb;
  return tc::launch<128, tc::EPI_BIAS_ACT>(a, lda, bt, ldb, g, st);
}
// split-K partials: part[z][M][ldp] = A[M, Kz] * Bt[N, Kz]^T ; the caller reduces over z in fixed order
int hrb_tc_splits(int64_t M, int32_t N, int32_t K) {
  const bool pair = tc::pair_enabled();
  const int bmt = pair ? 2 * tc::BM : tc::BM;
  const int tiles = (int)((M + bmt - 1) / bmt) * ((N + 127) / 128);
  const int kb_total = (K + tc::BK - 1) / tc::BK;
  int splits = (pair ? sm_count() / 2 : sm_count()) / (tiles > 0 ? tiles : 1);
  if (splits < 1) splits = 1;
  if (splits > kb_total) splits = kb_total;
  const int per = (kb_total + splits - 1) / splits;
  return (kb_total + per - 1) / per;
}
int hrb_tc_gemm_splitk(const float* a, int64_t lda, const float* bt, int64_t ldb, int64_t M, int32_t N, int32_t K, int32_t splits,
                       float* part, int64_t ldp, float* bsum, cudaStream_t st) {
  if (!(aligned16(a) && aligned16(bt) && lda % 4 == 0 && ldb % 4 == 0 && aligned16(part) && ldp % 4 == 0 && K >= 64))
    return fail(HRB_UNSUPPORTED, "tcgen05 split-K GEMM: shape/alignment not covered");
  const int kb_total = (K + tc::BK - 1) / tc::BK;
  const int per = (kb_total + splits - 1) / splits;
  if ((kb_total + per - 1) / per != splits) return fail(HRB_BAD_ARG, "tcgen05 split-K GEMM: %d splits leave empty slices", splits);
  tc::Args g{part, nullptr, nullptr, nullptr, ldp, 0, 0, M, N, K, 0, splits, per, nullptr, nullptr, 0, bsum, nullptr, 0};
  return tc::launch<128, tc::EPI_PLAIN>(a, lda, bt, ldb, g, st);
}

// same with the A operand untransposed: x[K samples][M features] row-major (the kernel transposes on its way into tensor memory)
int hrb_tc_gemm_splitk_an(const float* x, int64_t ldx, const float* bt, int64_t ldb, int64_t M, int32_t N, int32_t K, int32_t splits,
                          float* part, int64_t ldp, float* bsum, cudaStream_t st) {
  if (!(aligned16(x) && aligned16(bt) && ldx % 4 == 0 && ldb % 4 == 0 && aligned16(part) && ldp % 4 == 0 && K >= 64))
    return fail(HRB_UNSUPPORTED, "tcgen05 split-K GEMM: shape/alignment not covered");
  const int kb_total = (K + tc::BK - 1) / tc::BK;
  const int per = (kb_total + splits - 1) / splits;
  if ((kb_total + per - 1) / per != splits) return fail(HRB_BAD_ARG, "tcgen05 split-K GEMM: %d splits leave empty slices", splits);
  tc::Args g{part, nullptr, nullptr, nullptr, ldp, 0, 0, M, N, K, 0, splits, per, nullptr, nullptr, 0, bsum, nullptr, 0};
  return tc::launch<128, tc::EPI_PLAIN, true>(x, ldx, bt, ldb, g, st);
}

// ---- hooks used by dense.cu (row-major weight layouts of hrb_dense_*) ---------------------------------------
// The Dense entry points take W[K,N] (N contiguous).  Only bwd_x has both operands K-major as stored; fwd needs
// W^T and bwd_w needs x^T / dz^T, which the engine keeps as transposed copies and feeds to hrb_gemm_tn_* directly.
int hrb_tc_dense_fwd(const float*, int64_t, const float*, int64_t, const float*, int64_t, int32_t, int32_t, int32_t, float*, int64_t,
                     cudaStream_t) {
  return hrb::fail(HRB_UNSUPPORTED, "tcgen05 forward needs the transposed weight: call hrb_gemm_tn_bias_act");
}
int hrb_tc_dense_bwd_x(const float* dz, int64_t lddz, const float* w, int64_t ldw, int64_t M, int32_t K, int32_t N, const float* a_prev,
                       int64_t lda_prev, int32_t act_prev, float* dx, int64_t lddx, cudaStream_t st) {
  // dx[M,K] = dz[M,N] * w[K,N]^T: A = dz (N contiguous), Bt = w (K rows, N contiguous) -> exactly the TN form
  if (!tc::tc_ok(dz, lddz, w, ldw, M, K, N)) return hrb::fail(HRB_UNSUPPORTED, "tcgen05 bwd_x: shape/alignment not covered");
  if (a_prev != nullptr && !(aligned16(a_prev) && lda_prev % 4 == 0)) return hrb::fail(HRB_UNSUPPORTED, "tcgen05 bwd_x: a_prev alignment");
  if (!(aligned16(dx) && lddx % 4 == 0)) return hrb::fail(HRB_UNSUPPORTED, "tcgen05 bwd_x: dx alignment");
  tc::Args g{dx, nullptr, nullptr, a_prev, lddx, 0, lda_prev, M, K, N, act_prev, 1, (N + tc::BK - 1) / tc::BK, nullptr, nullptr, 0, nullptr, nullptr, 0};
  return tc::launch<128, tc::EPI_ACT_GRAD>(dz, lddz, w, ldw, g, st);
}
int hrb_tc_dense_bwd_w(const float*, int64_t, const float*, int64_t, int64_t, int32_t, int32_t, float*, int64_t, void*, size_t,
                       cudaStream_t) {
  return hrb::fail(HRB_UNSUPPORTED, "tcgen05 bwd_w needs transposed activations: call hrb_gemm_tn_splitk");
}
