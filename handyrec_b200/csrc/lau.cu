// lau.cu -- DIN local-activation attention, fused forward.
//
// Reference: layers/sequence.py:92-102 (LocalActivationUnit.call), layers/tools.py:104-113
// (SqueezeMask), models/ranking/sequential/DIN.py:87-93 (gather, attention, tf.matmul(att, keys)).
//   att_in[b,t,:] = [q, k, q-k, q*k]   (never materialised in HBM: built per tile in shared memory)
//   score[b,t]    = MLP(att_in)[0] * (key_id != 0)        -- no softmax
//   pooled[b,:]   = sum_t score[b,t] * k[b,t,:]
// MLP = Dense(4D) -> act -> Dense(h1) -> act -> ... -> Dense(1)   (core.py:57 prepends Dense(in))
// One CTA owns RT flattened (b,t) rows; activations ping-pong between two shared-memory tiles,
// weights stream through L1/L2 (<= 100 KB for D = 32), 4x4 register tiles.
#include "common.cuh"

namespace hrb {

constexpr int LAU_RT = 64;
constexpr int LAU_MAX_LAYERS = 6;
constexpr int HRB_ACT_DICE_ = 4;

struct LauLayers {
  int32_t n_layers;
  int32_t in[LAU_MAX_LAYERS];
  int32_t out[LAU_MAX_LAYERS];
  int64_t w_off[LAU_MAX_LAYERS];     // float offsets into params
  int64_t b_off[LAU_MAX_LAYERS];
  int64_t dice_off[LAU_MAX_LAYERS];  // alpha | mean | var (3*out) when act == dice
};

__global__ void __launch_bounds__(256) lau_mlp_kernel(const float* __restrict__ table, int64_t vocab, int32_t D,
                                                     const int32_t* __restrict__ qid, const int32_t* __restrict__ kid,
                                                     int64_t n_rows, int32_t T, const float* __restrict__ params,
                                                     LauLayers L, int32_t act, int32_t pitch, float* __restrict__ score) {
  extern __shared__ __align__(16) float smem[];
  float* X0 = smem;
  float* X1 = smem + (size_t)LAU_RT * pitch;
  const int tid = threadIdx.x;
  const int D4 = D / 4;
  for (int64_t tile = blockIdx.x; tile * LAU_RT < n_rows; tile += gridDim.x) {
    const int64_t r0 = tile * LAU_RT;
    // ---- build att_in: one float4 of (q,k) per item -> 4 float4 stores --------------------------
    for (int i = tid; i < LAU_RT * D4; i += 256) {
      const int r = i / D4, c = i - r * D4;
      const int64_t row = r0 + r;
      float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f), k4 = q4;
      if (row < n_rows) {
        const int64_t b = row / T;
        const int32_t qi = __ldg(qid + b), ki = __ldg(kid + row);
        if (qi >= 0 && qi < vocab) q4 = __ldg(reinterpret_cast<const float4*>(table + (int64_t)qi * D) + c);
        if (ki >= 0 && ki < vocab) k4 = __ldg(reinterpret_cast<const float4*>(table + (int64_t)ki * D) + c);
      }
      float* xr = X0 + (size_t)r * pitch + c * 4;
      *reinterpret_cast<float4*>(xr) = q4;
      *reinterpret_cast<float4*>(xr + D) = k4;
      *reinterpret_cast<float4*>(xr + 2 * D) = make_float4(q4.x - k4.x, q4.y - k4.y, q4.z - k4.z, q4.w - k4.w);
      *reinterpret_cast<float4*>(xr + 3 * D) = make_float4(q4.x * k4.x, q4.y * k4.y, q4.z * k4.z, q4.w * k4.w);
    }
    __syncthreads();
    float* Xin = X0;
    float* Xout = X1;
    for (int l = 0; l < L.n_layers; ++l) {
      const int K = L.in[l], N = L.out[l];
      const float* __restrict__ W = params + L.w_off[l];
      const float* __restrict__ bias = params + L.b_off[l];
      const bool last = (l == L.n_layers - 1);
      if (N % 4 == 0) {
        const int cg = N / 4;
        for (int item = tid; item < (LAU_RT / 4) * cg; item += 256) {
          const int rg = item / cg, c = item - rg * cg;
          float acc[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
          const float* xr = Xin + (size_t)(rg * 4) * pitch;
#pragma unroll 4
          for (int k = 0; k < K; ++k) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(W + (int64_t)k * N) + c);
            const float x0 = xr[k], x1 = xr[pitch + k], x2 = xr[2 * pitch + k], x3 = xr[3 * pitch + k];
            acc[0][0] = fmaf(x0, w4.x, acc[0][0]); acc[0][1] = fmaf(x0, w4.y, acc[0][1]);
            acc[0][2] = fmaf(x0, w4.z, acc[0][2]); acc[0][3] = fmaf(x0, w4.w, acc[0][3]);
            acc[1][0] = fmaf(x1, w4.x, acc[1][0]); acc[1][1] = fmaf(x1, w4.y, acc[1][1]);
            acc[1][2] = fmaf(x1, w4.z, acc[1][2]); acc[1][3] = fmaf(x1, w4.w, acc[1][3]);
            acc[2][0] = fmaf(x2, w4.x, acc[2][0]); acc[2][1] = fmaf(x2, w4.y, acc[2][1]);
            acc[2][2] = fmaf(x2, w4.z, acc[2][2]); acc[2][3] = fmaf(x2, w4.w, acc[2][3]);
            acc[3][0] = fmaf(x3, w4.x, acc[3][0]); acc[3][1] = fmaf(x3, w4.y, acc[3][1]);
            acc[3][2] = fmaf(x3, w4.z, acc[3][2]); acc[3][3] = fmaf(x3, w4.w, acc[3][3]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = c * 4 + j;
              float v = acc[i][j] + __ldg(bias + n);
              if (!last) {
                if (act == HRB_ACT_DICE_) {
                  const float* dp = params + L.dice_off[l];
                  const float p = sigmoidf_((v - __ldg(dp + N + n)) * rsqrtf(__ldg(dp + 2 * N + n) + 1e-9f));
                  v = p * v + (1.f - p) * __ldg(dp + n) * v;
                } else {
                  v = act_apply(act, v);
                }
              }
              Xout[(size_t)(rg * 4 + i) * pitch + n] = v;
            }
          }
        }
      } else {  // narrow layer (the final Dense(1)): one thread per (row, n)
        for (int item = tid; item < LAU_RT * N; item += 256) {
          const int r = item / N, n = item - r * N;
          float acc = 0.f;
          const float* xr = Xin + (size_t)r * pitch;
          for (int k = 0; k < K; ++k) acc = fmaf(xr[k], __ldg(W + (int64_t)k * N + n), acc);
          float v = acc + __ldg(bias + n);
          if (!last) {
            if (act == HRB_ACT_DICE_) {
              const float* dp = params + L.dice_off[l];
              const float p = sigmoidf_((v - __ldg(dp + N + n)) * rsqrtf(__ldg(dp + 2 * N + n) + 1e-9f));
              v = p * v + (1.f - p) * __ldg(dp + n) * v;
            } else {
              v = act_apply(act, v);
            }
          }
          Xout[(size_t)r * pitch + n] = v;
        }
      }
      __syncthreads();
      float* t = Xin;
      Xin = Xout;
      Xout = t;
    }
    // Xin now holds the (RT, 1) scores in column 0
    for (int r = tid; r < LAU_RT; r += 256) {
      const int64_t row = r0 + r;
      if (row < n_rows) score[row] = (__ldg(kid + row) != 0) ? Xin[(size_t)r * pitch] : 0.0f;  // sequence.py:101
    }
    __syncthreads();
  }
}

// pooled[b,:] = sum_t score[b,t] * table[kid[b,t],:]   (DIN.py:93)
__global__ void __launch_bounds__(256) lau_pool_kernel(const float* __restrict__ table, int64_t vocab, int32_t D4,
                                                      const int32_t* __restrict__ kid, const float* __restrict__ score,
                                                      int64_t batch, int32_t T, float* __restrict__ pooled) {
  const int64_t total = batch * D4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / D4;
    const int c = (int)(i - b * D4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < T; ++t) {
      const float s = __ldg(score + b * T + t);
      const int32_t ki = __ldg(kid + b * T + t);
      if (s != 0.f && ki >= 0 && ki < vocab) {
        const float4 k4 = __ldg(reinterpret_cast<const float4*>(table + (int64_t)ki * D4 * 4) + c);
        acc.x = fmaf(s, k4.x, acc.x); acc.y = fmaf(s, k4.y, acc.y);
        acc.z = fmaf(s, k4.z, acc.z); acc.w = fmaf(s, k4.w, acc.w);
      }
    }
    reinterpret_cast<float4*>(pooled)[i] = acc;
  }
}

}  // namespace hrb

using namespace hrb;

HRB_API int hrb_lau_fwd(const float* table, int64_t vocab, int32_t dim, const int32_t* query_ids, const int32_t* key_ids,
                        int64_t batch, int32_t seq_len, const float* params, const int32_t* layer_out_host,
                        int32_t n_layers, int32_t act, float* score, float* pooled, void* stream) {
  HRB_REQUIRE(table && query_ids && key_ids && params && layer_out_host && score && batch >= 0 && seq_len > 0 && dim > 0,
              "hrb_lau_fwd: null/negative argument");
  if (dim % 4 != 0) return fail(HRB_UNSUPPORTED, "hrb_lau_fwd: dim %d is not a multiple of 4", dim);
  HRB_REQUIRE(n_layers >= 1 && n_layers <= LAU_MAX_LAYERS, "hrb_lau_fwd: n_layers must be in [1,%d]", LAU_MAX_LAYERS);
  HRB_REQUIRE(layer_out_host[n_layers - 1] == 1, "hrb_lau_fwd: the last layer must have 1 unit (scores are (B,1,T))");
  HRB_REQUIRE(act >= HRB_ACT_LINEAR && act <= HRB_ACT_DICE_, "hrb_lau_fwd: unknown activation %d", act);
  HRB_REQUIRE(aligned16(table) && aligned16(params) && (pooled == nullptr || aligned16(pooled)),
              "hrb_lau_fwd: table/params/pooled must be 16-byte aligned");
  if (batch == 0) return HRB_OK;
  LauLayers L{};
  L.n_layers = n_layers;
  int in = 4 * dim, maxw = 4 * dim;
  int64_t off = 0;
  for (int l = 0; l < n_layers; ++l) {
    const int out = layer_out_host[l];
    HRB_REQUIRE(out > 0, "hrb_lau_fwd: layer %d has %d units", l, out);
    L.in[l] = in;
    L.out[l] = out;
    L.w_off[l] = off;
    off += (int64_t)in * out;
    L.b_off[l] = off;
    off += out;
    L.dice_off[l] = off;
    if (act == HRB_ACT_DICE_ && l != n_layers - 1) off += 3 * (int64_t)out;
    off = (off + 3) / 4 * 4;  // keep every W 16-byte aligned
    if (out > maxw) maxw = out;
    in = out;
  }
  const int pitch = maxw + 4;
  const size_t smem = (size_t)2 * LAU_RT * pitch * sizeof(float);
  if (smem > 200 * 1024) return fail(HRB_UNSUPPORTED, "hrb_lau_fwd: layer width %d too large for the fused kernel", maxw);
  HRB_CUDA(cudaFuncSetAttribute(lau_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_rows = batch * seq_len;
  int64_t tiles = (n_rows + LAU_RT - 1) / LAU_RT;
  const int per_sm = (int)((220 * 1024) / (smem + 1024));
  const int64_t cap = (int64_t)sm_count() * (per_sm < 1 ? 1 : per_sm);
  if (tiles > cap) tiles = cap;
  lau_mlp_kernel<<<(unsigned)tiles, 256, smem, st>>>(table, vocab, dim, query_ids, key_ids, n_rows, seq_len, params, L, act,
                                                    pitch, score);
  HRB_LAUNCH_CHECK();
  if (pooled != nullptr) {
    const int64_t items = batch * (dim / 4);
    int64_t blocks = (items + 255) / 256;
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
    lau_pool_kernel<<<(unsigned)blocks, 256, 0, st>>>(table, vocab, dim / 4, key_ids, score, batch, seq_len, pooled);
    HRB_LAUNCH_CHECK();
  }
  return HRB_OK;
}

// =============================================================================================
// a10, TRAINING: the first layer of the local-activation MLP without the (B,T,4D) tensor.
//   att_in = [q, k, q-k, q*k] (sequence.py:96-97) and W = [Wa; Wb; Wc; Wd] (4 blocks of D rows) give
//     att_in . W = q.(Wa + Wc) + k.(Wb - Wc) + (q*k).Wd
//   i.e. a per-SAMPLE term  qterm[b] = q[b].(Wa + Wc) + bias  (B x U, tiny)  plus a B*T-row GEMM over  A' = [k | q*k]  (2D columns
//   instead of 4D: half the flops, half the bytes), added in the GEMM epilogue as a bias per group of T rows
//   (hrb_dense_fwd_t_grouped).  The kernels here are the element-wise glue around the three GEMMs: weight split / gradient merge,
//   A' and its backward, the per-sample sum of dz.  All reductions run in fixed order.
// =============================================================================================
namespace hrb {

__global__ void __launch_bounds__(256) lau_split_weights_kernel(const float* __restrict__ W, int64_t ldw, int32_t D, int32_t U,
                                                               float* __restrict__ wq, float* __restrict__ wp, float* __restrict__ wpt) {
  const int total = D * U;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int d = i / U, u = i - d * U;
    const float a = W[(int64_t)d * ldw + u], b = W[(int64_t)(D + d) * ldw + u], c = W[(int64_t)(2 * D + d) * ldw + u],
                e = W[(int64_t)(3 * D + d) * ldw + u];
    wq[i] = a + c;
    wp[i] = b - c;
    wp[(int64_t)(D + d) * U + u] = e;
    wpt[(int64_t)u * 2 * D + d] = b - c;
    wpt[(int64_t)u * 2 * D + D + d] = e;
  }
}

__global__ void __launch_bounds__(256) lau_merge_wgrads_kernel(const float* __restrict__ dwq, const float* __restrict__ dwp, int32_t D, int32_t U,
                                                              float* __restrict__ dW, int64_t ldw) {
  const int total = D * U;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int d = i / U, u = i - d * U;
    const float gq = dwq[i], gk = dwp[i], gm = dwp[(int64_t)(D + d) * U + u];
    dW[(int64_t)d * ldw + u] = gq;             // Wa: through the q term
    dW[(int64_t)(D + d) * ldw + u] = gk;       // Wb: through the k term
    dW[(int64_t)(2 * D + d) * ldw + u] = gq - gk;  // Wc multiplies (q - k)
    dW[(int64_t)(3 * D + d) * ldw + u] = gm;   // Wd: through q*k
  }
}

// A'[m, :D] = k[m, :],  A'[m, D:] = q[m / T, :] * k[m, :]   (float4 granularity, D % 4 == 0)
__global__ void __launch_bounds__(256) lau_pack_fwd_kernel(const float* __restrict__ q, const float* __restrict__ k, int64_t rows, int32_t T,
                                                          int32_t D4, float* __restrict__ out) {
  const int64_t total = rows * D4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / D4;
    const int c = (int)(i - m * D4);
    const float4 kv = __ldg(reinterpret_cast<const float4*>(k) + i);
    const float4 qv = __ldg(reinterpret_cast<const float4*>(q) + (m / T) * D4 + c);
    float4* o = reinterpret_cast<float4*>(out) + m * 2 * D4;
    o[c] = kv;
    o[D4 + c] = make_float4(qv.x * kv.x, qv.y * kv.y, qv.z * kv.z, qv.w * kv.w);
  }
}

// dk[m] = dA[m, :D] + dA[m, D:] * q[b];   dq[b] (+)= sum_t dA[m, D:] * k[m]   (one thread per (sample, float4 column): fixed order over t)
__global__ void __launch_bounds__(256) lau_pack_bwd_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ dA,
                                                          int64_t batch, int32_t T, int32_t D4, int32_t accumulate, float* __restrict__ dq,
                                                          float* __restrict__ dk) {
  const int64_t total = batch * D4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / D4;
    const int c = (int)(i - b * D4);
    const float4 qv = __ldg(reinterpret_cast<const float4*>(q) + i);
    float4 acc = accumulate ? reinterpret_cast<const float4*>(dq)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = 0; t < T; ++t) {
      const int64_t m = b * T + t;
      const float4 gk = __ldg(reinterpret_cast<const float4*>(dA) + m * 2 * D4 + c);
      const float4 gm = __ldg(reinterpret_cast<const float4*>(dA) + m * 2 * D4 + D4 + c);
      const float4 kv = __ldg(reinterpret_cast<const float4*>(k) + m * D4 + c);
      reinterpret_cast<float4*>(dk)[m * D4 + c] = make_float4(gk.x + gm.x * qv.x, gk.y + gm.y * qv.y, gk.z + gm.z * qv.z, gk.w + gm.w * qv.w);
      acc.x = fmaf(gm.x, kv.x, acc.x); acc.y = fmaf(gm.y, kv.y, acc.y); acc.z = fmaf(gm.z, kv.z, acc.z); acc.w = fmaf(gm.w, kv.w, acc.w);
    }
    reinterpret_cast<float4*>(dq)[i] = acc;
  }
}

// out[g, :] = sum_{t < T} x[g*T + t, :]
__global__ void __launch_bounds__(256) group_sum_kernel(const float* __restrict__ x, int64_t ldx, int64_t groups, int32_t T, int32_t N,
                                                       float* __restrict__ out) {
  const int64_t total = groups * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = i / N;
    const int n = (int)(i - g * N);
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc += __ldg(x + (g * T + t) * ldx + n);
    out[i] = acc;
  }
}

static inline unsigned lgrid(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace hrb

HRB_API int hrb_lau_split_weights(const float* w, int64_t ldw, int32_t dim, int32_t units, float* wq, float* wp, float* wpt, void* stream) {
  HRB_REQUIRE(w && wq && wp && wpt && dim > 0 && units > 0 && ldw >= units, "hrb_lau_split_weights: bad argument");
  hrb::lau_split_weights_kernel<<<hrb::lgrid((int64_t)dim * units), 256, 0, (cudaStream_t)stream>>>(w, ldw, dim, units, wq, wp, wpt);
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}

HRB_API int hrb_lau_merge_wgrads(const float* dwq, const float* dwp, int32_t dim, int32_t units, float* dw, int64_t ldw, void* stream) {
  HRB_REQUIRE(dwq && dwp && dw && dim > 0 && units > 0 && ldw >= units, "hrb_lau_merge_wgrads: bad argument");
  hrb::lau_merge_wgrads_kernel<<<hrb::lgrid((int64_t)dim * units), 256, 0, (cudaStream_t)stream>>>(dwq, dwp, dim, units, dw, ldw);
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}

HRB_API int hrb_lau_pack_fwd(const float* q, const float* k, int64_t batch, int32_t T, int32_t dim, float* out, void* stream) {
  HRB_REQUIRE(q && k && out && batch >= 0 && T > 0 && dim > 0, "hrb_lau_pack_fwd: bad argument");
  if (dim % 4 != 0) return hrb::fail(HRB_UNSUPPORTED, "hrb_lau_pack_fwd: dim %d is not a multiple of 4", dim);
  HRB_REQUIRE(hrb::aligned16(q) && hrb::aligned16(k) && hrb::aligned16(out), "hrb_lau_pack_fwd: buffers must be 16-byte aligned");
  if (batch == 0) return HRB_OK;
  hrb::lau_pack_fwd_kernel<<<hrb::lgrid(batch * T * (dim / 4)), 256, 0, (cudaStream_t)stream>>>(q, k, batch * T, T, dim / 4, out);
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}

HRB_API int hrb_lau_pack_bwd(const float* q, const float* k, const float* da, int64_t batch, int32_t T, int32_t dim, int32_t accumulate_dq,
                             float* dq, float* dk, void* stream) {
  HRB_REQUIRE(q && k && da && dq && dk && batch >= 0 && T > 0 && dim > 0, "hrb_lau_pack_bwd: bad argument");
  if (dim % 4 != 0) return hrb::fail(HRB_UNSUPPORTED, "hrb_lau_pack_bwd: dim %d is not a multiple of 4", dim);
  HRB_REQUIRE(hrb::aligned16(q) && hrb::aligned16(k) && hrb::aligned16(da) && hrb::aligned16(dq) && hrb::aligned16(dk),
              "hrb_lau_pack_bwd: buffers must be 16-byte aligned");
  if (batch == 0) return HRB_OK;
  hrb::lau_pack_bwd_kernel<<<hrb::lgrid(batch * (dim / 4)), 256, 0, (cudaStream_t)stream>>>(q, k, da, batch, T, dim / 4, accumulate_dq, dq, dk);
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}

HRB_API int hrb_group_sum(const float* x, int64_t ldx, int64_t groups, int32_t group_rows, int32_t n, float* out, void* stream) {
  HRB_REQUIRE(x && out && groups >= 0 && group_rows > 0 && n > 0 && ldx >= n, "hrb_group_sum: bad argument");
  if (groups == 0) return HRB_OK;
  hrb::group_sum_kernel<<<hrb::lgrid(groups * n), 256, 0, (cudaStream_t)stream>>>(x, ldx, groups, group_rows, n, out);
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}
