// retrieval.cu -- the loss side of the retrieval models: tf.nn.sampled_softmax_loss as used by
// SampledSoftmaxLayer (layers/tools.py:32-84) and tf.nn.l2_normalize as used by DSSM (models/retrieval/DSSM.py:105-106).
//
// sampled_softmax_loss (TF 2.x defaults: num_true = 1, remove_accidental_hits = True, subtract_log_q through
// _compute_sampled_logits, log-uniform candidate sampler with unique = True):
//   z_0   = <u_b, w_label(b)> - log Q(label(b))                 (true class)
//   z_j   = <u_b, w_sampled(j)> - log Q(sampled(j))             j = 1..S, -FLT_MAX where sampled(j) == label(b)
//   loss_b = logsumexp(z) - z_0
// with Q(k) = 1 - (1 - p_k)^num_tries, p_k = log((k+2)/(k+1)) / log(range_max + 1): the expected count of class k in a batch of
// `num_sampled` unique draws that took `num_tries` tries.  The dot products are separate GEMM / row-dot kernels; this file
// holds the fused correction + mask + softmax cross-entropy (forward and backward in one pass over the (B, 1+S) logits).
#include <float.h>

#include "common.cuh"

namespace hrb {

__device__ __forceinline__ float log_expected_count(int32_t k, float num_tries, float inv_log_range) {
  const float p = logf(((float)k + 2.0f) / ((float)k + 1.0f)) * inv_log_range;
  return logf(-expm1f(num_tries * log1pf(-p)));
}

// one warp per row
__global__ void __launch_bounds__(256) sampled_softmax_kernel(const float* __restrict__ true_logit, const float* __restrict__ samp_logit,
                                                             int64_t ld, const int32_t* __restrict__ labels,
                                                             const int32_t* __restrict__ sampled, int64_t batch, int32_t S,
                                                             const float* __restrict__ true_expected, const float* __restrict__ samp_expected,
                                                             float num_tries, float inv_log_range, int32_t remove_hits,
                                                             const float* __restrict__ gout, float* __restrict__ loss,
                                                             float* __restrict__ d_true, float* __restrict__ d_samp) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= batch) return;
  const int32_t lab = labels[b];
  const float z0 = true_logit[b] - (true_expected != nullptr ? logf(true_expected[b]) : log_expected_count(lab, num_tries, inv_log_range));
  const float* row = samp_logit + b * ld;
  float mx = z0;
  for (int j = lane; j < S; j += 32) {
    const int32_t c = sampled[j];
    float z = row[j] - (samp_expected != nullptr ? logf(samp_expected[j]) : log_expected_count(c, num_tries, inv_log_range));
    if (remove_hits && c == lab) z = -FLT_MAX;
    mx = fmaxf(mx, z);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = lane == 0 ? expf(z0 - mx) : 0.f;
  for (int j = lane; j < S; j += 32) {
    const int32_t c = sampled[j];
    float z = row[j] - (samp_expected != nullptr ? logf(samp_expected[j]) : log_expected_count(c, num_tries, inv_log_range));
    if (remove_hits && c == lab) z = -FLT_MAX;
    sum += expf(z - mx);
  }
  sum = warp_sum(sum);
  const float lse = mx + logf(sum);
  if (lane == 0 && loss != nullptr) loss[b] = lse - z0;
  if (d_samp == nullptr) return;
  const float g = gout != nullptr ? gout[b] : 1.0f;
  if (lane == 0) d_true[b] = g * (expf(z0 - lse) - 1.0f);
  float* drow = d_samp + b * ld;
  for (int j = lane; j < S; j += 32) {
    const int32_t c = sampled[j];
    float z = row[j] - (samp_expected != nullptr ? logf(samp_expected[j]) : log_expected_count(c, num_tries, inv_log_range));
    const bool hit = remove_hits && c == lab;
    drow[j] = hit ? 0.f : g * expf(z - lse);
  }
}

// out[m] = <a[m,:], b[m,:]>; one warp per row
__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb,
                                                    int64_t M, int32_t K, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (m >= M) return;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s = fmaf(a[m * lda + k], b[m * ldb + k], s);
  s = warp_sum(s);
  if (lane == 0) out[m] = s;
}

// out[m,:] = s[m] * x[m,:]
__global__ void __launch_bounds__(256) rowscale_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ s, int64_t M,
                                                      int32_t K, float* __restrict__ out, int64_t ldo) {
  const int64_t total = M * K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / K;
    const int k = (int)(i - m * K);
    out[m * ldo + k] = s[m] * x[m * ldx + k];
  }
}

// fixed-order two-stage reduction of sum_i a[i]*b[i] (b == nullptr: a[i]^2): partial[block] then one CTA adds them in order
constexpr int RED_BLOCKS = 256;
__global__ void __launch_bounds__(256) dot_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                         float* __restrict__ partial) {
  __shared__ float red[8];
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = blockIdx.x * per, hi = min(n, lo + per);
  float s = 0.f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) s = fmaf(a[i], b != nullptr ? b[i] : a[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k];
    partial[blockIdx.x] = t;
  }
}
// mode 0 (forward): stat[0] = rsqrt(max(sum, eps)); y = x * stat[0]
// mode 1 (backward): dot = sum(dy*y); dx = (dy - y*dot) * stat[0]
__global__ void __launch_bounds__(256) l2norm_apply_kernel(const float* __restrict__ partial, int mode, float eps, float* __restrict__ stat,
                                                          const float* __restrict__ x, const float* __restrict__ dy, int64_t n,
                                                          float* __restrict__ out) {
  float t = 0.f;
  for (int k = 0; k < RED_BLOCKS; ++k) t += partial[k];  // every thread adds the partials in the same order
  if (mode == 0) {
    const float inv = rsqrtf(fmaxf(t, eps));
    if (blockIdx.x == 0 && threadIdx.x == 0) stat[0] = inv;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = x[i] * inv;
  } else {
    const float inv = stat[0];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
      out[i] = (dy[i] - x[i] * t) * inv;  // x = y here
  }
}

static inline unsigned rgrid(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace hrb

using namespace hrb;

HRB_API int hrb_sampled_softmax(const float* true_logit, const float* sampled_logit, int64_t ld, const int32_t* labels,
                                const int32_t* sampled, int64_t batch, int32_t num_sampled, const float* true_expected,
                                const float* sampled_expected, float num_tries, int64_t range_max, int32_t remove_accidental_hits,
                                const float* gout, float* loss, float* d_true_logit, float* d_sampled_logit, void* stream) {
  HRB_REQUIRE(true_logit && sampled_logit && labels && sampled && batch >= 0 && num_sampled > 0 && ld >= num_sampled && range_max > 0,
              "hrb_sampled_softmax: bad argument");
  HRB_REQUIRE((d_true_logit == nullptr) == (d_sampled_logit == nullptr), "hrb_sampled_softmax: give both gradient buffers or none");
  HRB_REQUIRE((true_expected == nullptr) == (sampled_expected == nullptr), "hrb_sampled_softmax: give both expected counts or none");
  if (batch == 0) return HRB_OK;
  const float inv_log_range = 1.0f / logf((float)range_max + 1.0f);
  sampled_softmax_kernel<<<(unsigned)((batch + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      true_logit, sampled_logit, ld, labels, sampled, batch, num_sampled, true_expected, sampled_expected, num_tries, inv_log_range,
      remove_accidental_hits, gout, loss, d_true_logit, d_sampled_logit);
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}

HRB_API int hrb_rowdot(const float* a, int64_t lda, const float* b, int64_t ldb, int64_t M, int32_t K, float* out, void* stream) {
  HRB_REQUIRE(a && b && out && M >= 0 && K > 0, "hrb_rowdot: bad argument");
  if (M == 0) return HRB_OK;
  rowdot_kernel<<<(unsigned)((M + 7) / 8), 256, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, M, K, out);
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}

HRB_API int hrb_rowscale(const float* x, int64_t ldx, const float* s, int64_t M, int32_t K, float* out, int64_t ldo, void* stream) {
  HRB_REQUIRE(x && s && out && M >= 0 && K > 0, "hrb_rowscale: bad argument");
  if (M == 0) return HRB_OK;
  rowscale_kernel<<<rgrid(M * K), 256, 0, (cudaStream_t)stream>>>(x, ldx, s, M, K, out, ldo);
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}

HRB_API int hrb_l2_normalize_fwd(const float* x, int64_t n, float eps, float* y, float* stat, float* scratch, void* stream) {
  HRB_REQUIRE(x && y && stat && scratch && n >= 0, "hrb_l2_normalize_fwd: bad argument");
  if (n == 0) return HRB_OK;
  dot_partial_kernel<<<RED_BLOCKS, 256, 0, (cudaStream_t)stream>>>(x, nullptr, n, scratch);
  HRB_LAUNCH_CHECK();
  l2norm_apply_kernel<<<rgrid(n), 256, 0, (cudaStream_t)stream>>>(scratch, 0, eps, stat, x, nullptr, n, y);
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}

HRB_API int hrb_l2_normalize_bwd(const float* y, const float* dy, int64_t n, const float* stat, float* dx, float* scratch, void* stream) {
  HRB_REQUIRE(y && dy && stat && dx && scratch && n >= 0, "hrb_l2_normalize_bwd: bad argument");
  if (n == 0) return HRB_OK;
  dot_partial_kernel<<<RED_BLOCKS, 256, 0, (cudaStream_t)stream>>>(dy, y, n, scratch);
  HRB_LAUNCH_CHECK();
  l2norm_apply_kernel<<<rgrid(n), 256, 0, (cudaStream_t)stream>>>(scratch, 1, 0.f, const_cast<float*>(stat), y, dy, n, dx);
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}
