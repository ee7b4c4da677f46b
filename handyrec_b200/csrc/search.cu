// search.cu -- top-n inner-product search over the item embeddings: the step after the retrieval towers, done by
// faiss.IndexFlatIP in the reference (handyrec/models/utils.py:7-51: index.add(item_embd); index.search(user_embd, n)).
//
// The scores of a chunk of items are one GEMM (queries . items^T, the tcgen05 kernel for Q >= 512); this kernel folds a
// (Q, C) chunk of scores into the running top-n of every query.  One CTA per query: the n-th best score so far is a threshold,
// only candidates above it are collected (after the first chunks almost none), candidates + current list are ordered by a
// bitonic sort on (score desc, index asc) in shared memory -- the result is the exact top n with faiss' tie order (lower id
// first) and does not depend on the order the candidates were collected in.
#include <float.h>

#include "common.cuh"

namespace hrb {

constexpr int TK_THREADS = 256;
constexpr int TK_SUB = 1024;   // columns scanned per round: candidates (<= TK_SUB) + list (<= 1024) fit the 2048-slot sort buffer
constexpr int TK_SLOTS = 2048;

__device__ __forceinline__ bool tk_before(float va, int ia, float vb, int ib) { return va > vb || (va == vb && ia < ib); }

__global__ void __launch_bounds__(TK_THREADS) topk_merge_kernel(const float* __restrict__ scores, int64_t ld, int32_t C, int64_t col_base,
                                                                int32_t k, float* __restrict__ best_val, int32_t* __restrict__ best_idx) {
  __shared__ float sv[TK_SLOTS];
  __shared__ int si[TK_SLOTS];
  __shared__ int n_cand;
  const int tid = threadIdx.x;
  const int64_t q = blockIdx.x;
  const float* row = scores + q * ld;
  float* bv = best_val + q * k;
  int32_t* bi = best_idx + q * k;
  for (int c0 = 0; c0 < C; c0 += TK_SUB) {
    const int c1 = min(C, c0 + TK_SUB);
    const float thr_v = bv[k - 1];
    const int thr_i = bi[k - 1];
    if (tid == 0) n_cand = 0;
    __syncthreads();
    for (int c = c0 + tid; c < c1; c += TK_THREADS) {
      const float v = row[c];
      const int idx = (int)(col_base + c);
      if (thr_i < 0 || tk_before(v, idx, thr_v, thr_i)) {  // thr_i < 0: the list is not full yet
        const int slot = atomicAdd(&n_cand, 1);             // integer bookkeeping; the sort below fixes the order
        sv[slot] = v;
        si[slot] = idx;
      }
    }
    __syncthreads();
    const int nc = n_cand;
    if (nc == 0) continue;  // uniform
    for (int j = tid; j < k; j += TK_THREADS) {
      sv[nc + j] = bv[j];
      si[nc + j] = bi[j];
    }
    int n = 1;
    while (n < nc + k) n <<= 1;
    for (int j = nc + k + tid; j < n; j += TK_THREADS) {
      sv[j] = -FLT_MAX;
      si[j] = -1;
    }
    __syncthreads();
    for (int size = 2; size <= n; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = tid; t < (n >> 1); t += TK_THREADS) {
          const int lo = 2 * t - (t & (stride - 1));
          const int hi = lo + stride;
          const bool up = (lo & size) == 0;  // "up" blocks end with the best element first
          const float va = sv[lo], vb = sv[hi];
          const int ia = si[lo], ib = si[hi];
          // order so that empty slots (index -1) sink to the end
          const bool a_first = ib < 0 ? true : (ia < 0 ? false : tk_before(va, ia, vb, ib));
          if (a_first != up) {
            sv[lo] = vb; sv[hi] = va;
            si[lo] = ib; si[hi] = ia;
          }
        }
        __syncthreads();
      }
    }
    for (int j = tid; j < k; j += TK_THREADS) {
      bv[j] = sv[j];
      bi[j] = si[j];
    }
    __syncthreads();
  }
}

__global__ void topk_init_kernel(float* __restrict__ best_val, int32_t* __restrict__ best_idx, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    best_val[i] = -FLT_MAX;
    best_idx[i] = -1;
  }
}

}  // namespace hrb

using namespace hrb;

HRB_API int hrb_topk_init(float* best_val, int32_t* best_idx, int64_t queries, int32_t k, void* stream) {
  HRB_REQUIRE(best_val && best_idx && queries >= 0 && k > 0, "hrb_topk_init: bad argument");
  if (queries == 0) return HRB_OK;
  const int64_t n = queries * k;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  topk_init_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(best_val, best_idx, n);
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}

HRB_API int hrb_topk_merge(const float* scores, int64_t ld, int64_t queries, int32_t n_cols, int64_t col_base, int32_t k, float* best_val,
                           int32_t* best_idx, void* stream) {
  HRB_REQUIRE(scores && best_val && best_idx && queries >= 0 && n_cols >= 0 && ld >= n_cols && col_base >= 0, "hrb_topk_merge: bad argument");
  if (k <= 0 || k > 1024) return fail(HRB_UNSUPPORTED, "hrb_topk_merge: k = %d outside [1, 1024]", k);
  HRB_REQUIRE(col_base + n_cols <= 0x7fffffffll, "hrb_topk_merge: item index exceeds 2^31-1");
  if (queries == 0 || n_cols == 0) return HRB_OK;
  topk_merge_kernel<<<(unsigned)queries, TK_THREADS, 0, (cudaStream_t)stream>>>(scores, ld, n_cols, col_base, k, best_val, best_idx);
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}
