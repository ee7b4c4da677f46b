// fm.cu -- FM second-order interaction on a materialised (B,F,D) tensor.
//
// Reference: /root/reference/handyrec/layers/interaction.py:15-39
//   part2 = sum_f Dense1(x_f) = (sum_f x_f) . w          (w in R^{D x 1}, no bias)
//   part3 = 0.5 * sum_d [ (sum_f x)^2 - sum_f x^2 ]
//   out   = part2 + part3 + w0
//
// Forward: the pooled field embeddings of a tile of samples are staged in shared memory with
// 1-D bulk TMA copies (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP), double buffered, one
// copy per sample row into a padded row pitch so that the float4 reads of the G = D/4 lanes that own
// a sample are bank-conflict free.  HBM-bound: F*D*4 bytes read + 4 (+ D*4 for fm_sum) written / sample.
#include "common.cuh"

namespace hrb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a lost transaction becomes a trap (-> CUDA error), never a hung GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 4000000000LL) __trap();  // ~2 s
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// G lanes per sample, TS = blockDim/G samples per tile, 2 stages.
template <int G>
__global__ void __launch_bounds__(256) fm_fwd_tma_kernel(const float* __restrict__ x, int64_t x_ld, int64_t batch,
                                                        int32_t F, int32_t pitch /* floats */,
                                                        const float* __restrict__ w, const float* __restrict__ w0,
                                                        float* __restrict__ out, float* __restrict__ fm_sum) {
  constexpr int D = G * 4;
  constexpr int TS = 256 / G;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stage[2];
  stage[0] = reinterpret_cast<float*>(smem_raw);
  stage[1] = stage[0] + (size_t)TS * pitch;
  __shared__ __align__(8) uint64_t full[2];

  const int q = threadIdx.x % G;
  const int s = threadIdx.x / G;
  const uint32_t row_bytes = (uint32_t)(F * D * 4);
  const int64_t n_tiles = (batch + TS - 1) / TS;

  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int64_t tile, int st) {  // executed by thread 0 only
    const int64_t b0 = tile * TS;
    const int rows = (int)min((int64_t)TS, batch - b0);
    mbar_expect_tx(&full[st], row_bytes * (uint32_t)rows);
    for (int r = 0; r < rows; ++r)
      bulk_g2s(stage[st] + (size_t)r * pitch, x + (b0 + r) * x_ld, row_bytes, &full[st]);
  };

  const float4 w4 = __ldg(reinterpret_cast<const float4*>(w) + q);
  const float bias = __ldg(w0);
  int64_t tile = blockIdx.x;
  if (tile < n_tiles && threadIdx.x == 0) issue(tile, 0);
  uint32_t phase[2] = {0, 0};
  int st = 0;
  for (; tile < n_tiles; tile += gridDim.x, st ^= 1) {
    const int64_t next = tile + gridDim.x;
    if (next < n_tiles && threadIdx.x == 0) issue(next, st ^ 1);  // stage st^1 was drained one iteration ago
    mbar_wait(&full[st], phase[st]);
    phase[st] ^= 1;
    const int64_t b = tile * TS + s;
    const bool valid = b < batch;  // lanes past the end still take part in the shuffles below
    float4 S = make_float4(0.f, 0.f, 0.f, 0.f), Q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
      const float4* row = reinterpret_cast<const float4*>(stage[st] + (size_t)s * pitch) + q;
#pragma unroll 4
      for (int f = 0; f < F; ++f) {
        const float4 v = row[f * G];
        S.x += v.x; S.y += v.y; S.z += v.z; S.w += v.w;
        Q.x = fmaf(v.x, v.x, Q.x); Q.y = fmaf(v.y, v.y, Q.y); Q.z = fmaf(v.z, v.z, Q.z); Q.w = fmaf(v.w, v.w, Q.w);
      }
    }
    float p2 = S.x * w4.x + S.y * w4.y + S.z * w4.z + S.w * w4.w;
    float p3 = (S.x * S.x - Q.x) + (S.y * S.y - Q.y) + (S.z * S.z - Q.z) + (S.w * S.w - Q.w);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      p2 += __shfl_xor_sync(0xffffffffu, p2, o, G);
      p3 += __shfl_xor_sync(0xffffffffu, p3, o, G);
    }
    if (valid) {
      if (q == 0) out[b] = p2 + 0.5f * p3 + bias;
      if (fm_sum != nullptr) reinterpret_cast<float4*>(fm_sum + b * D)[q] = S;
    }
    __syncthreads();  // every lane is done with stage st before it is refilled two tiles later
  }
}

// fallback for shapes the staged kernel does not cover (D % 4 != 0, very wide rows): one warp per sample
__global__ void __launch_bounds__(256) fm_fwd_direct_kernel(const float* __restrict__ x, int64_t x_ld, int64_t batch,
                                                           int32_t F, int32_t D, const float* __restrict__ w,
                                                           const float* __restrict__ w0, float* __restrict__ out,
                                                           float* __restrict__ fm_sum) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < batch; b += n_warps) {
    float p2 = 0.f, p3 = 0.f;
    for (int d = lane; d < D; d += 32) {
      float S = 0.f, Q = 0.f;
      for (int f = 0; f < F; ++f) {
        const float v = x[b * x_ld + (int64_t)f * D + d];
        S += v;
        Q = fmaf(v, v, Q);
      }
      p2 = fmaf(S, w[d], p2);
      p3 += S * S - Q;
      if (fm_sum != nullptr) fm_sum[b * D + d] = S;
    }
    p2 = warp_sum(p2);
    p3 = warp_sum(p3);
    if (lane == 0) out[b] = p2 + 0.5f * p3 + w0[0];
  }
}

// backward: dx[b,f,:] (+)= dout[b] * (w + S[b] - x[b,f,:]);  dw += dout[b]*S[b];  dw0 += dout[b]
template <int G>
__global__ void __launch_bounds__(256) fm_bwd_kernel(const float* __restrict__ x, int64_t x_ld, int64_t batch, int32_t F,
                                                    const float* __restrict__ w, const float* __restrict__ dout,
                                                    float* __restrict__ dx, int64_t dx_ld, int32_t accumulate,
                                                    float* __restrict__ dw_dw0) {
  constexpr int D = G * 4;
  constexpr int TS = 256 / G;
  __shared__ float red[256 / 32][D + 1];
  const int q = threadIdx.x % G;
  const int s = threadIdx.x / G;
  const float4 w4 = __ldg(reinterpret_cast<const float4*>(w) + q);
  float4 gw = make_float4(0.f, 0.f, 0.f, 0.f);
  float gw0 = 0.f;
  for (int64_t b = blockIdx.x * (int64_t)TS + s; b < batch; b += (int64_t)gridDim.x * TS) {
    const float g = __ldg(dout + b);
    const float4* xr = reinterpret_cast<const float4*>(x + b * x_ld) + q;
    float4* dr = reinterpret_cast<float4*>(dx + b * dx_ld) + q;
    float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int FU = 8;
    // pass 1: S = sum_f x
    for (int f0 = 0; f0 < F; f0 += FU) {
      float4 v[FU];
#pragma unroll
      for (int u = 0; u < FU; ++u) v[u] = (f0 + u < F) ? __ldg(xr + (f0 + u) * G) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < FU; ++u) { S.x += v[u].x; S.y += v[u].y; S.z += v[u].z; S.w += v[u].w; }
    }
    const float4 base = make_float4(g * (w4.x + S.x), g * (w4.y + S.y), g * (w4.z + S.z), g * (w4.w + S.w));
    // pass 2 (rows are L1/L2 hot): dx = base - g*x
    for (int f0 = 0; f0 < F; f0 += FU) {
      float4 v[FU], o[FU];
#pragma unroll
      for (int u = 0; u < FU; ++u) {
        v[u] = (f0 + u < F) ? __ldg(xr + (f0 + u) * G) : make_float4(0.f, 0.f, 0.f, 0.f);
        o[u] = (accumulate && f0 + u < F) ? dr[(f0 + u) * G] : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < FU; ++u) {
        if (f0 + u < F) {
          o[u].x += base.x - g * v[u].x; o[u].y += base.y - g * v[u].y;
          o[u].z += base.z - g * v[u].z; o[u].w += base.w - g * v[u].w;
          dr[(f0 + u) * G] = o[u];
        }
      }
    }
    gw.x = fmaf(g, S.x, gw.x); gw.y = fmaf(g, S.y, gw.y); gw.z = fmaf(g, S.z, gw.z); gw.w = fmaf(g, S.w, gw.w);
    if (q == 0) gw0 += g;
  }
  // block reduction over the TS sample slots: lanes with equal q (stride G inside a warp), then warps
#pragma unroll
  for (int o = 16; o >= G; o >>= 1) {
    gw.x += __shfl_xor_sync(0xffffffffu, gw.x, o);
    gw.y += __shfl_xor_sync(0xffffffffu, gw.y, o);
    gw.z += __shfl_xor_sync(0xffffffffu, gw.z, o);
    gw.w += __shfl_xor_sync(0xffffffffu, gw.w, o);
    gw0 += __shfl_xor_sync(0xffffffffu, gw0, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (G <= 32 && lane < G) {
    red[warp][lane * 4 + 0] = gw.x; red[warp][lane * 4 + 1] = gw.y;
    red[warp][lane * 4 + 2] = gw.z; red[warp][lane * 4 + 3] = gw.w;
    if (lane == 0) red[warp][D] = gw0;
  }
  __syncthreads();
  if (threadIdx.x <= D) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < 256 / 32; ++wv) t += red[wv][threadIdx.x];
    atomicAdd(dw_dw0 + threadIdx.x, t);  // one atomic per CTA per parameter (D+1 scalars)
  }
}

static inline unsigned capped_grid(int64_t blocks, int per_sm) {
  const int64_t cap = (int64_t)sm_count() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace hrb

using namespace hrb;

HRB_API int hrb_fm_fwd(const float* x, int64_t x_ld, int64_t batch, int32_t fields, int32_t dim, const float* w,
                       const float* w0, float* out, float* fm_sum, void* stream) {
  HRB_REQUIRE(x && w && w0 && out && batch >= 0 && fields > 0 && dim > 0, "hrb_fm_fwd: null/negative argument");
  HRB_REQUIRE(x_ld >= (int64_t)fields * dim, "hrb_fm_fwd: x_ld %lld < fields*dim", (long long)x_ld);
  if (batch == 0) return HRB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int G = dim / 4;
  const bool pow2 = dim % 4 == 0 && (G & (G - 1)) == 0 && G <= 32;
  const bool aligned = aligned16(x) && aligned16(w) && x_ld % 4 == 0 && (fm_sum == nullptr || aligned16(fm_sum));
  if (pow2 && aligned) {
    const int TS = 256 / G;
    // pad the row pitch so consecutive samples start 4G banks apart (conflict-free LDS.128 for G < 8)
    int pitch = fields * dim;
    const int want = (4 * G) % 32;
    while (pitch % 32 != want) pitch += 4;
    const size_t smem = (size_t)2 * TS * pitch * sizeof(float);
    if (smem <= 200 * 1024) {
#define HRB_FM_FWD(GG)                                                                                             \
  {                                                                                                                \
    static bool attr_set = false;                                                                                  \
    if (!attr_set) {                                                                                               \
      HRB_CUDA(cudaFuncSetAttribute(fm_fwd_tma_kernel<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
      attr_set = true;                                                                                             \
    }                                                                                                              \
    const int per_sm = smem <= 100 * 1024 ? 2 : 1;                                                                 \
    fm_fwd_tma_kernel<GG><<<capped_grid((batch + TS - 1) / TS, per_sm), 256, smem, st>>>(x, x_ld, batch, fields, pitch, \
                                                                                          w, w0, out, fm_sum);    \
  }
      switch (G) {
        case 1: HRB_FM_FWD(1) break;
        case 2: HRB_FM_FWD(2) break;
        case 4: HRB_FM_FWD(4) break;
        case 8: HRB_FM_FWD(8) break;
        case 16: HRB_FM_FWD(16) break;
        default: HRB_FM_FWD(32) break;
      }
#undef HRB_FM_FWD
      HRB_LAUNCH_CHECK();
      return HRB_OK;
    }
  }
  fm_fwd_direct_kernel<<<capped_grid((batch + 7) / 8, 8), 256, 0, st>>>(x, x_ld, batch, fields, dim, w, w0, out, fm_sum);
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}

HRB_API int hrb_fm_bwd(const float* x, int64_t x_ld, int64_t batch, int32_t fields, int32_t dim, const float* w,
                       const float* dout, float* dx, int64_t dx_ld, int32_t accumulate, float* dw_dw0, void* stream) {
  HRB_REQUIRE(x && w && dout && dx && dw_dw0 && batch >= 0 && fields > 0 && dim > 0, "hrb_fm_bwd: null/negative argument");
  const int G = dim / 4;
  if (dim % 4 != 0 || (G & (G - 1)) != 0 || G > 32)
    return fail(HRB_UNSUPPORTED, "hrb_fm_bwd: dim/4 must be a power of two <= 32 (dim=%d)", dim);
  HRB_REQUIRE(aligned16(x) && aligned16(dx) && aligned16(w) && x_ld % 4 == 0 && dx_ld % 4 == 0,
              "hrb_fm_bwd: x/dx/w must be 16-byte aligned with leading dims %% 4 == 0");
  cudaStream_t st = (cudaStream_t)stream;
  HRB_CUDA(cudaMemsetAsync(dw_dw0, 0, sizeof(float) * (dim + 1), st));
  if (batch == 0) return HRB_OK;
  const int TS = 256 / G;
  const unsigned grid = capped_grid((batch + TS - 1) / TS, 8);
  switch (G) {
    case 1: fm_bwd_kernel<1><<<grid, 256, 0, st>>>(x, x_ld, batch, fields, w, dout, dx, dx_ld, accumulate, dw_dw0); break;
    case 2: fm_bwd_kernel<2><<<grid, 256, 0, st>>>(x, x_ld, batch, fields, w, dout, dx, dx_ld, accumulate, dw_dw0); break;
    case 4: fm_bwd_kernel<4><<<grid, 256, 0, st>>>(x, x_ld, batch, fields, w, dout, dx, dx_ld, accumulate, dw_dw0); break;
    case 8: fm_bwd_kernel<8><<<grid, 256, 0, st>>>(x, x_ld, batch, fields, w, dout, dx, dx_ld, accumulate, dw_dw0); break;
    case 16: fm_bwd_kernel<16><<<grid, 256, 0, st>>>(x, x_ld, batch, fields, w, dout, dx, dx_ld, accumulate, dw_dw0); break;
    default: fm_bwd_kernel<32><<<grid, 256, 0, st>>>(x, x_ld, batch, fields, w, dout, dx, dx_ld, accumulate, dw_dw0); break;
  }
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}
