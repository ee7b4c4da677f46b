// dense.cu -- Dense layers of the reference DNN (layers/core.py:53-78; Keras Dense = x@W+b),
// forward and both backward GEMMs, fp32 FFMA path.
//
// One tiled SGEMM template serves the three products (row-major everywhere):
//   fwd    y [M,N] = act( x[M,K]  @ w[K,N] + bias )        A = x   (k contiguous), B = w  (n contiguous)
//   bwd_x  dx[M,K] = (dz[M,N] @ w[K,N]^T) * act'(a_prev)   A = dz  (k contiguous), B = w^T (k contiguous)
//   bwd_w  dw[K,N] =  x[M,K]^T @ dz[M,N]                   A = x^T (m contiguous), B = dz (n contiguous), split over M
// 128x128x16 CTA tile, 256 threads, 8x8 register tile, register-prefetch double buffering.
// Bound: fp32 FFMA pipe (148 SMs x 128 FMA/clk).  The tcgen05 3xTF32 path lives in gemm_tc.cu.
#include "common.cuh"

namespace hrb {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8;
constexpr int PAD = 4;

enum Epilogue { EPI_BIAS_ACT = 0, EPI_ACT_GRAD = 1, EPI_PLAIN = 2 };

struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  int64_t lda, ldb, ldc;
  int64_t M;     // rows of C
  int32_t N;     // cols of C
  int64_t Kred;  // reduction length
  // epilogue
  const float* bias;   // [N]           (EPI_BIAS_ACT)
  const float* aprev;  // [M, ldaprev]  (EPI_ACT_GRAD)
  int64_t ldaprev;
  int32_t act;
  // split over the reduction dim (bwd_w): slice z handles [z*kslice, min(Kred,(z+1)*kslice)) and writes C + z*M*ldc
  int64_t kslice;
};

// load a 4-vector with per-element bounds masking along the contiguous dim
__device__ __forceinline__ float4 load4(const float* __restrict__ p, int64_t idx_contig, int64_t extent, bool row_ok,
                                        bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!row_ok || idx_contig >= extent) return v;
  if (vec && idx_contig + 3 < extent) return __ldg(reinterpret_cast<const float4*>(p));
  v.x = __ldg(p);
  if (idx_contig + 1 < extent) v.y = __ldg(p + 1);
  if (idx_contig + 2 < extent) v.z = __ldg(p + 2);
  if (idx_contig + 3 < extent) v.w = __ldg(p + 3);
  return v;
}

// TA: A is stored [k][m] (m contiguous);  !TA: A is stored [m][k] (k contiguous)
// TB: B is stored [n][k] (k contiguous);  !TB: B is stored [k][n] (n contiguous)
template <bool TA, bool TB, int EPI>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g, bool vecA, bool vecB, bool vecC) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * g.kslice;
  const int64_t kend = min(g.Kred, kbeg + g.kslice);
  float* __restrict__ C = g.C + (int64_t)blockIdx.z * g.M * g.ldc;

  const int ty = tid / 16, tx = tid % 16;  // 16x16 threads, each 8x8 outputs (rows ty*8.., cols tx*8..)
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto gload = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (!TA) {  // 128 rows x 16 k: thread -> (m = tid/4 + 64 i, k4 = (tid%4)*4)
        const int m = tid / 4 + 64 * i, k4 = (tid % 4) * 4;
        const int64_t gm = m0 + m, gk = k0 + k4;
        ra[i] = load4(g.A + gm * g.lda + gk, gk, kend, gm < g.M, vecA);
      } else {    // 16 k x 128 m: thread -> (k = tid/32 + 8 i, m4 = (tid%32)*4)
        const int k = tid / 32 + 8 * i, m4 = (tid % 32) * 4;
        const int64_t gm = m0 + m4, gk = k0 + k;
        ra[i] = load4(g.A + gk * g.lda + gm, gm, g.M, gk < kend, vecA);
      }
      if (!TB) {  // 16 k x 128 n
        const int k = tid / 32 + 8 * i, n4 = (tid % 32) * 4;
        const int64_t gn = n0 + n4, gk = k0 + k;
        rb[i] = load4(g.B + gk * g.ldb + gn, gn, g.N, gk < kend, vecB);
      } else {    // 128 n x 16 k
        const int n = tid / 4 + 64 * i, k4 = (tid % 4) * 4;
        const int64_t gn = n0 + n, gk = k0 + k4;
        rb[i] = load4(g.B + gn * g.ldb + gk, gk, kend, gn < g.N, vecB);
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (!TA) {
        const int m = tid / 4 + 64 * i, k4 = (tid % 4) * 4;
        As[buf][k4 + 0][m] = ra[i].x; As[buf][k4 + 1][m] = ra[i].y;
        As[buf][k4 + 2][m] = ra[i].z; As[buf][k4 + 3][m] = ra[i].w;
      } else {
        const int k = tid / 32 + 8 * i, m4 = (tid % 32) * 4;
        *reinterpret_cast<float4*>(&As[buf][k][m4]) = ra[i];
      }
      if (!TB) {
        const int k = tid / 32 + 8 * i, n4 = (tid % 32) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][k][n4]) = rb[i];
      } else {
        const int n = tid / 4 + 64 * i, k4 = (tid % 4) * 4;
        Bs[buf][k4 + 0][n] = rb[i].x; Bs[buf][k4 + 1][n] = rb[i].y;
        Bs[buf][k4 + 2][n] = rb[i].z; Bs[buf][k4 + 3][n] = rb[i].w;
      }
    }
  };

  int buf = 0;
  if (kbeg < kend) {
    gload(kbeg);
    sstore(0);
  }
  __syncthreads();
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = k0 + BK < kend;
    if (more) gload(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      // split 8 = 4 + 4 with a 64-column stride between halves keeps LDS.128 conflict-free
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4 + 64]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4 + 64]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }

  // epilogue: rows {ty*4..+3, 64+ty*4..+3}, cols {tx*4..+3, 64+tx*4..+3}
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= g.M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int gn = n0 + jh * 64 + tx * 4;
      if (gn >= g.N) continue;
      float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (gn + c < g.N) {
          if (EPI == EPI_BIAS_ACT) {
            const float bz = g.bias != nullptr ? __ldg(g.bias + gn + c) : 0.f;
            v[c] = act_apply(g.act, v[c] + bz);
          } else if (EPI == EPI_ACT_GRAD) {
            if (g.aprev != nullptr) v[c] *= act_grad_from_out(g.act, __ldg(g.aprev + gm * g.ldaprev + gn + c));
          }
        }
      }
      float* cp = C + gm * g.ldc + gn;
      if (vecC && gn + 3 < g.N) {
        *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (gn + c < g.N) cp[c] = v[c];
      }
    }
  }
}

// fixed-order reduction of the split partials: out[i] = sum_z part[z, i]
__global__ void __launch_bounds__(256) split_reduce_kernel(const float* __restrict__ part, int64_t rows, int32_t cols,
                                                          int64_t ld_part, int32_t splits, float* __restrict__ out,
                                                          int64_t ld_out) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[((int64_t)z * rows + r) * ld_part + c];
    out[r * ld_out + c] = s;
  }
}

// many partials per output (split counts in the hundreds): one warp per output element, lanes stride over the
// partials (independent loads in flight), fixed-shape tree at the end -> still deterministic
__global__ void __launch_bounds__(256) split_reduce_wide_kernel(const float* __restrict__ part, int64_t rows, int32_t cols,
                                                               int64_t ld_part, int32_t splits, float* __restrict__ out,
                                                               int64_t ld_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t total = rows * cols;
  for (int64_t i = warp0; i < total; i += nwarps) {
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    float s = 0.f;
    for (int z = lane; z < splits; z += 32) s += __ldg(part + ((int64_t)z * rows + r) * ld_part + c);
    s = warp_sum(s);
    if (lane == 0) out[r * ld_out + c] = s;
  }
}

static inline void launch_split_reduce(const float* part, int64_t rows, int32_t cols, int64_t ld_part, int32_t splits, float* out,
                                       int64_t ld_out, cudaStream_t st) {
  const int64_t total = rows * cols;
  if (splits > 16 && total * 32 <= (int64_t)sm_count() * 2048 * 4) {
    const int64_t blocks = min((int64_t)sm_count() * 8, (total * 32 + 255) / 256);
    split_reduce_wide_kernel<<<(unsigned)blocks, 256, 0, st>>>(part, rows, cols, ld_part, splits, out, ld_out);
  } else {
    const int64_t blocks = min((int64_t)sm_count() * 4, (total + 255) / 256);
    split_reduce_kernel<<<(unsigned)(blocks < 1 ? 1 : blocks), 256, 0, st>>>(part, rows, cols, ld_part, splits, out, ld_out);
  }
}

// Two reductions with the same split count in ONE launch (weight-gradient partials + bias-gradient partials of a layer): the first
// a.blocks CTAs work on job a, the rest on job b, each with the scheme launch_split_reduce would have picked for it alone, so every
// output element is summed in exactly the same order as by two separate launches.
struct ReduceJob {
  const float* part;
  int64_t rows;
  int32_t cols;
  int64_t ld_part;
  float* out;
  int64_t ld_out;
  int32_t wide;
  int32_t blocks;
};
__global__ void __launch_bounds__(256) split_reduce2_kernel(const ReduceJob a, const ReduceJob b, int32_t splits) {
  const bool first = (int)blockIdx.x < a.blocks;
  const ReduceJob& j = first ? a : b;
  const int64_t bid = first ? blockIdx.x : blockIdx.x - a.blocks;
  const int64_t total = j.rows * j.cols;
  if (!j.wide) {
    for (int64_t i = bid * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)j.blocks * blockDim.x) {
      const int64_t r = i / j.cols;
      const int c = (int)(i - r * j.cols);
      float s = 0.f;
      for (int z = 0; z < splits; ++z) s += j.part[((int64_t)z * j.rows + r) * j.ld_part + c];
      j.out[r * j.ld_out + c] = s;
    }
    return;
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (bid * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)j.blocks * blockDim.x) >> 5;
  for (int64_t i = warp0; i < total; i += nwarps) {
    const int64_t r = i / j.cols;
    const int c = (int)(i - r * j.cols);
    float s = 0.f;
    for (int z = lane; z < splits; z += 32) s += __ldg(j.part + ((int64_t)z * j.rows + r) * j.ld_part + c);
    s = warp_sum(s);
    if (lane == 0) j.out[r * j.ld_out + c] = s;
  }
}
static inline ReduceJob reduce_job(const float* part, int64_t rows, int32_t cols, int64_t ld_part, int32_t splits, float* out, int64_t ld_out) {
  const int64_t total = rows * cols;
  ReduceJob j{part, rows, cols, ld_part, out, ld_out, 0, 1};
  if (splits > 16 && total * 32 <= (int64_t)sm_count() * 2048 * 4) {
    j.wide = 1;
    j.blocks = (int32_t)min((int64_t)sm_count() * 8, (total * 32 + 255) / 256);
  } else {
    j.blocks = (int32_t)max((int64_t)1, min((int64_t)sm_count() * 4, (total + 255) / 256));
  }
  return j;
}
static inline void launch_split_reduce2(const ReduceJob& a, const ReduceJob& b, int32_t splits, cudaStream_t st) {
  split_reduce2_kernel<<<(unsigned)(a.blocks + b.blocks), 256, 0, st>>>(a, b, splits);
}

// column sums of dz[M,N] (bias gradient), two fixed-order passes: per-CTA partials then a final sum
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ dz, int64_t lddz, int64_t M,
                                                            int32_t N, int64_t rows_per_block,
                                                            float* __restrict__ part /* [gridDim.y, N] */) {
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;  // 8 row lanes
  __shared__ float red[8][33];
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  if (n < N)
    for (int64_t r = r0 + ry; r < r1; r += 8) s += __ldg(dz + r * lddz + n);
  red[ry][threadIdx.x & 31] = s;
  __syncthreads();
  if (ry == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x & 31];
    part[(int64_t)blockIdx.y * N + n] = t;
  }
}

__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, int64_t n,
                                                     int32_t act, float* __restrict__ dz) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dz[i] = dy[i] * act_grad_from_out(act, y[i]);
}

// ---- the 1-unit logit layer (DeepFM.py:59-60 forces units[-1] == 1): GEMV forward, fused backward ----------
// y[m] = x[m,:] . w + b ; one warp per row, float4 loads
__global__ void __launch_bounds__(256) dense1_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                                                        const float* __restrict__ bias, int64_t M, int32_t K, float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float b0 = bias != nullptr ? __ldg(bias) : 0.f;
  for (int64_t m = warp0; m < M; m += nwarps) {
    const float* xr = x + m * ldx;
    float acc = 0.f;
    for (int k = lane * 4; k < K; k += 128) {
      if (k + 3 < K) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(xr + k));
        const float4 b = __ldg(reinterpret_cast<const float4*>(w + k));
        acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc); acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
      } else {
        for (int j = k; j < K; ++j) acc = fmaf(__ldg(xr + j), __ldg(w + j), acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) y[m] = acc + b0;
  }
}

// Backward of the 1-unit layer as three streaming kernels:
//   (1) dz_prev[m,k] = dy[m] * w[k] * act'(x[m,k])            element-wise, float4
//   (2) dz_prev^T via hrb_transpose_kernel                      (only when the tensor-core bwd_w needs it)
//   (3) dw[k] = sum_m x[m,k]*dy[m], db = sum_m dy[m]            per-slab partials + fixed-order reduce
__global__ void __launch_bounds__(256) dense1_bwd_dz_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                                                           const float* __restrict__ dy, int64_t M, int32_t K4 /* K/4 */, int32_t act,
                                                           float* __restrict__ dzp, int64_t lddz) {
  const int64_t total = M * K4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / K4;
    const int c = (int)(i - m * K4);
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + m * ldx) + c);
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + c);
    const float g = __ldg(dy + m);
    float4 o;
    o.x = g * wv.x * act_grad_from_out(act, xv.x);
    o.y = g * wv.y * act_grad_from_out(act, xv.y);
    o.z = g * wv.z * act_grad_from_out(act, xv.z);
    o.w = g * wv.w * act_grad_from_out(act, xv.w);
    reinterpret_cast<float4*>(dzp + m * lddz)[c] = o;
  }
}

// part[blockIdx.y][k] = sum over the row slab of x[m,k]*dy[m]; part[blockIdx.y][K] = sum dy  (32 columns x 8 row lanes)
// The logit layer's backward in one pass over x: a CTA walks 32-row tiles, writes dz, stages x and dz in shared memory, writes the
// transposed dz tile (128-byte rows of dz^T) and keeps its dw / dbias partial sums in registers across tiles (fixed tile and row
// order -> deterministic).  Shared memory: 2 x 32 x (K+1) floats, so K <= 190; wider layers take the three-kernel path.
constexpr int D1_ROWS = 32;
__global__ void __launch_bounds__(256) dense1_bwd_fused_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w,
                                                              const float* __restrict__ dy, int64_t M, int32_t K, int32_t act,
                                                              float* __restrict__ dzp, int64_t lddz, float* __restrict__ dzt, int64_t lddzt,
                                                              float* __restrict__ part) {
  extern __shared__ float d1_smem[];
  const int P = K + 1, K4 = K >> 2;
  float* xs = d1_smem;
  float* zs = xs + D1_ROWS * P;
  float* dys = zs + D1_ROWS * P;
  const int64_t n_tiles = (M + D1_ROWS - 1) / D1_ROWS;
  float acc = 0.f, accb = 0.f;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t m0 = tile * D1_ROWS;
    if (threadIdx.x < D1_ROWS) dys[threadIdx.x] = (m0 + threadIdx.x < M) ? __ldg(dy + m0 + threadIdx.x) : 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < D1_ROWS * K4; i += 256) {
      const int r = i / K4, c = i - r * K4;
      const int64_t m = m0 + r;
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M) {
        xv = __ldg(reinterpret_cast<const float4*>(x + m * ldx) + c);
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + c);
        const float g = dys[r];
        o.x = g * wv.x * act_grad_from_out(act, xv.x);
        o.y = g * wv.y * act_grad_from_out(act, xv.y);
        o.z = g * wv.z * act_grad_from_out(act, xv.z);
        o.w = g * wv.w * act_grad_from_out(act, xv.w);
        reinterpret_cast<float4*>(dzp + m * lddz)[c] = o;
      }
      float* xr = xs + r * P + 4 * c;
      float* zr = zs + r * P + 4 * c;
      xr[0] = xv.x; xr[1] = xv.y; xr[2] = xv.z; xr[3] = xv.w;
      zr[0] = o.x; zr[1] = o.y; zr[2] = o.z; zr[3] = o.w;
    }
    __syncthreads();
    if (dzt != nullptr) {
      for (int i = threadIdx.x; i < K * D1_ROWS; i += 256) {
        const int k = i / D1_ROWS, r = i - k * D1_ROWS;
        if (m0 + r < M) dzt[(int64_t)k * lddzt + m0 + r] = zs[r * P + k];
      }
    }
    if ((int)threadIdx.x < K) {
#pragma unroll 8
      for (int r = 0; r < D1_ROWS; ++r) acc = fmaf(xs[r * P + threadIdx.x], dys[r], acc);
    }
    if (threadIdx.x == 255) {
#pragma unroll 8
      for (int r = 0; r < D1_ROWS; ++r) accb += dys[r];
    }
    __syncthreads();  // the next tile overwrites dys / xs / zs
  }
  if ((int)threadIdx.x < K) part[(int64_t)blockIdx.x * (K + 1) + threadIdx.x] = acc;
  if (threadIdx.x == 255) part[(int64_t)blockIdx.x * (K + 1) + K] = accb;
}
__global__ void __launch_bounds__(256) dense1_bwd_dw_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ dy,
                                                           int64_t M, int32_t K, int64_t rows_per_block, float* __restrict__ part) {
  const int k = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  __shared__ float red[8][33], redb[8];
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s = 0.f, sb = 0.f;
  for (int64_t m = r0 + ry; m < r1; m += 8) {
    const float g = __ldg(dy + m);
    if (k < K) s = fmaf(__ldg(x + m * ldx + k), g, s);
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) sb += g;
  }
  red[ry][threadIdx.x & 31] = s;
  if ((threadIdx.x & 31) == 0) redb[ry] = sb;
  __syncthreads();
  if (ry == 0) {
    if (k < K) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x & 31];
      part[(int64_t)blockIdx.y * (K + 1) + k] = t;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) t += redb[j];
      part[(int64_t)blockIdx.y * (K + 1) + K] = t;
    }
  }
}

static inline bool vec_ok(const void* p, int64_t ld) { return aligned16(p) && ld % 4 == 0; }

static int pick_splits(int64_t M, int32_t K, int32_t N) {
  const int tiles = ((K + BM - 1) / BM) * ((N + BN - 1) / BN);
  int splits = (2 * sm_count() + tiles - 1) / tiles;
  const int64_t max_by_len = (M + 4 * BK - 1) / (4 * BK);
  if (splits > max_by_len) splits = (int)max_by_len;
  if (splits < 1) splits = 1;
  if (splits > 256) splits = 256;
  return splits;
}

}  // namespace hrb

using namespace hrb;

template <bool TA, bool TB, int EPI>
static int launch_sgemm(const GemmArgs& g, int splits, cudaStream_t st) {
  dim3 grid((g.N + BN - 1) / BN, (unsigned)((g.M + BM - 1) / BM), splits);
  sgemm_kernel<TA, TB, EPI><<<grid, 256, 0, st>>>(g, vec_ok(g.A, g.lda), vec_ok(g.B, g.ldb), vec_ok(g.C, g.ldc));
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}

// 32x32 tiled transpose through shared memory: dst[c][r] = src[r][c]
__global__ void __launch_bounds__(256) hrb_transpose_kernel(const float* __restrict__ src, int64_t rows, int32_t cols, int64_t lds,
                                                        float* __restrict__ dst, int64_t ldd) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int64_t r0 = (int64_t)blockIdx.y * 32;
  const int c0 = blockIdx.x * 32;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = r0 + ty + 8 * i;
    tile[ty + 8 * i][tx] = (r < rows && c0 + tx < cols) ? src[r * lds + c0 + tx] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i;
    if (c < cols && r0 + tx < rows) dst[(int64_t)c * ldd + r0 + tx] = tile[tx][ty + 8 * i];
  }
}

// implemented in gemm_tc.cu (tcgen05 3xTF32); return HRB_UNSUPPORTED when the shape is not covered
int hrb_tc_gemm_bias_act(const float* a, int64_t lda, const float* bt, int64_t ldb, const float* bias, int64_t M, int32_t N, int32_t K,
                         int32_t act, float* c, int64_t ldc, float* ct, int64_t ldct, uint32_t* relu_mask, int64_t mask_ld, cudaStream_t st);
int hrb_tc_gemm_groupbias_act(const float* a, int64_t lda, const float* bt, int64_t ldb, const float* gbias, int64_t ldgb, int32_t group,
                              int64_t M, int32_t N, int32_t K, int32_t act, float* c, int64_t ldc, cudaStream_t st);
int hrb_tc_gemm_act_grad(const float* a, int64_t lda, const float* bt, int64_t ldb, int64_t M, int32_t N, int32_t K, const float* aprev,
                         int64_t ldap, int32_t act, float* c, int64_t ldc, float* ct, int64_t ldct, const uint32_t* relu_mask, int64_t mask_ld,
                         cudaStream_t st);
int hrb_tc_splits(int64_t M, int32_t N, int32_t K);
int hrb_tc_gemm_splitk(const float* a, int64_t lda, const float* bt, int64_t ldb, int64_t M, int32_t N, int32_t K, int32_t splits,
                       float* part, int64_t ldp, float* bsum, cudaStream_t st);
int hrb_tc_gemm_splitk_an(const float* x, int64_t ldx, const float* bt, int64_t ldb, int64_t M, int32_t N, int32_t K, int32_t splits,
                          float* part, int64_t ldp, float* bsum, cudaStream_t st);
int hrb_tc_dense_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, int64_t M, int32_t K,
                     int32_t N, int32_t act, float* y, int64_t ldy, cudaStream_t st);
int hrb_tc_dense_bwd_x(const float* dz, int64_t lddz, const float* w, int64_t ldw, int64_t M, int32_t K, int32_t N,
                       const float* a_prev, int64_t lda_prev, int32_t act_prev, float* dx, int64_t lddx, cudaStream_t st);
int hrb_tc_dense_bwd_w(const float* x, int64_t ldx, const float* dz, int64_t lddz, int64_t M, int32_t K, int32_t N,
                       float* dw, int64_t lddw, void* workspace, size_t workspace_bytes, cudaStream_t st);

HRB_API int hrb_dense_fwd(const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, int64_t M,
                          int32_t K, int32_t N, int32_t act, float* y, int64_t ldy, int32_t mode, void* stream) {
  HRB_REQUIRE(x && w && y && M >= 0 && K > 0 && N > 0 && ldx >= K && ldw >= N && ldy >= N, "hrb_dense_fwd: bad argument");
  HRB_REQUIRE(act >= HRB_ACT_LINEAR && act <= HRB_ACT_TANH, "hrb_dense_fwd: unknown activation %d", act);
  if (M == 0) return HRB_OK;
  if (mode != HRB_GEMM_FP32) {
    int rc = hrb_tc_dense_fwd(x, ldx, w, ldw, bias, M, K, N, act, y, ldy, (cudaStream_t)stream);
    if (rc == HRB_OK) return rc;
    if (mode == HRB_GEMM_3XTF32 || rc != HRB_UNSUPPORTED) return rc;
  }
  GemmArgs g{x, w, y, ldx, ldw, ldy, M, N, K, bias, nullptr, 0, act, K};
  return launch_sgemm<false, false, EPI_BIAS_ACT>(g, 1, (cudaStream_t)stream);
}

HRB_API int hrb_dense_bwd_x(const float* dz, int64_t lddz, const float* w, int64_t ldw, int64_t M, int32_t K, int32_t N,
                            const float* a_prev, int64_t lda_prev, int32_t act_prev, float* dx, int64_t lddx,
                            int32_t mode, void* stream) {
  HRB_REQUIRE(dz && w && dx && M >= 0 && K > 0 && N > 0 && lddz >= N && ldw >= N && lddx >= K, "hrb_dense_bwd_x: bad argument");
  HRB_REQUIRE(a_prev == nullptr || lda_prev >= K, "hrb_dense_bwd_x: lda_prev < K");
  if (M == 0) return HRB_OK;
  if (mode != HRB_GEMM_FP32) {
    int rc = hrb_tc_dense_bwd_x(dz, lddz, w, ldw, M, K, N, a_prev, lda_prev, act_prev, dx, lddx, (cudaStream_t)stream);
    if (rc == HRB_OK) return rc;
    if (mode == HRB_GEMM_3XTF32 || rc != HRB_UNSUPPORTED) return rc;
  }
  // dx[M,K] = dz[M,N] @ w[K,N]^T : C cols = K, reduction = N, B = w stored [K][N] = "[n][k]" with roles swapped
  GemmArgs g{dz, w, dx, lddz, ldw, lddx, M, K, N, nullptr, a_prev, lda_prev, act_prev, N};
  return launch_sgemm<false, true, EPI_ACT_GRAD>(g, 1, (cudaStream_t)stream);
}

HRB_API int hrb_dense_bwd_w_workspace(int64_t M, int32_t K, int32_t N, size_t* bytes) {
  HRB_REQUIRE(bytes && M >= 0 && K > 0 && N > 0, "hrb_dense_bwd_w_workspace: bad argument");
  const int splits = pick_splits(M, K, N);
  const size_t part = (size_t)splits * K * (size_t)((N + 3) / 4 * 4) * sizeof(float);
  const size_t colsum = (size_t)256 * N * sizeof(float);
  *bytes = part + colsum + 512;
  return HRB_OK;
}

HRB_API int hrb_dense_bwd_w(const float* x, int64_t ldx, const float* dz, int64_t lddz, int64_t M, int32_t K, int32_t N,
                            float* dw, int64_t lddw, float* dbias, void* workspace, size_t workspace_bytes, int32_t mode,
                            void* stream) {
  HRB_REQUIRE(x && dz && dw && workspace && M >= 0 && K > 0 && N > 0 && ldx >= K && lddz >= N && lddw >= N,
              "hrb_dense_bwd_w: bad argument");
  size_t need = 0;
  hrb_dense_bwd_w_workspace(M, K, N, &need);
  if (workspace_bytes < need) return fail(HRB_WORKSPACE, "hrb_dense_bwd_w: workspace %zu < required %zu bytes", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0) {
    HRB_CUDA(cudaMemset2DAsync(dw, lddw * sizeof(float), 0, N * sizeof(float), K, st));
    if (dbias) HRB_CUDA(cudaMemsetAsync(dbias, 0, N * sizeof(float), st));
    return HRB_OK;
  }
  const int splits = pick_splits(M, K, N);
  const int64_t ldp = (N + 3) / 4 * 4;
  float* part = (float*)workspace;
  float* colpart = part + (size_t)splits * K * ldp;
  bool done = false;
  if (mode != HRB_GEMM_FP32) {
    int rc = hrb_tc_dense_bwd_w(x, ldx, dz, lddz, M, K, N, dw, lddw, workspace, workspace_bytes, st);
    if (rc == HRB_OK) done = true;
    else if (mode == HRB_GEMM_3XTF32 || rc != HRB_UNSUPPORTED) return rc;
  }
  if (!done) {
    // C[K,N] = x^T @ dz : A = x stored [m][k] -> as A^T it is "[k_red][m_out]" = TA layout with lda = ldx
    int64_t kslice = (M + splits - 1) / splits;
    kslice = (kslice + BK - 1) / BK * BK;
    const int eff_splits = (int)((M + kslice - 1) / kslice);
    GemmArgs g{x, dz, part, ldx, lddz, ldp, K, N, M, nullptr, nullptr, 0, 0, kslice};
    int rc = launch_sgemm<true, false, EPI_PLAIN>(g, eff_splits, st);
    if (rc != HRB_OK) return rc;
    launch_split_reduce(part, K, N, ldp, eff_splits, dw, lddw, st);
    HRB_LAUNCH_CHECK();
  }
  if (dbias != nullptr) {
    int yb = (int)min((int64_t)256, (M + 255) / 256);
    if (yb < 1) yb = 1;
    const int64_t rpb = (M + yb - 1) / yb;
    dim3 grid((N + 31) / 32, yb);
    colsum_partial_kernel<<<grid, 256, 0, st>>>(dz, lddz, M, N, rpb, colpart);
    HRB_LAUNCH_CHECK();
    launch_split_reduce(colpart, 1, N, N, yb, dbias, N, st);
    HRB_LAUNCH_CHECK();
  }
  return HRB_OK;
}

HRB_API int hrb_act_bwd(const float* y, const float* dy, int64_t n, int32_t act, float* dz, void* stream) {
  HRB_REQUIRE(y && dy && dz && n >= 0, "hrb_act_bwd: bad argument");
  if (n == 0) return HRB_OK;
  const int64_t blocks = min((int64_t)sm_count() * 8, (n + 255) / 256);
  act_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(y, dy, n, act, dz);
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}

// ---------------------------------------------------------------------------------------------
// "TN" Dense entry points for the tcgen05 path: every operand has its reduction dim contiguous, so the
// caller keeps transposed copies (W^T for the forward, x^T and dz^T for the weight gradient).
// ---------------------------------------------------------------------------------------------
HRB_API int hrb_transpose(const float* src, int64_t rows, int32_t cols, int64_t lds, float* dst, int64_t ldd, void* stream) {
  HRB_REQUIRE(rows >= 0 && cols >= 0 && lds >= cols && ldd >= rows, "hrb_transpose: bad sizes");
  if (rows == 0 || cols == 0) return HRB_OK;
  HRB_REQUIRE(src && dst, "hrb_transpose: null pointer");
  dim3 grid((cols + 31) / 32, (unsigned)((rows + 31) / 32));
  HRB_REQUIRE(grid.y <= 65535u * 32u, "hrb_transpose: too many rows");
  if (grid.y > 65535) return fail(HRB_UNSUPPORTED, "hrb_transpose: rows > 2M");
  hrb_transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, rows, cols, lds, dst, ldd);
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}

HRB_API int hrb_dense_fwd_t(const float* x, int64_t ldx, const float* wt, int64_t ldwt, const float* bias, int64_t M, int32_t K,
                            int32_t N, int32_t act, float* y, int64_t ldy, float* yt, int64_t ldyt, uint32_t* relu_mask, void* stream) {
  HRB_REQUIRE(x && wt && y && M >= 0 && K > 0 && N > 0 && ldx >= K && ldwt >= K && ldy >= N, "hrb_dense_fwd_t: bad argument");
  HRB_REQUIRE(yt == nullptr || ldyt >= M, "hrb_dense_fwd_t: ldyt < M");
  HRB_REQUIRE(act >= HRB_ACT_LINEAR && act <= HRB_ACT_TANH, "hrb_dense_fwd_t: unknown activation %d", act);
  if (M == 0) return HRB_OK;
  return hrb_tc_gemm_bias_act(x, ldx, wt, ldwt, bias, M, N, K, act, y, ldy, yt, ldyt, relu_mask, (N + 31) / 32, (cudaStream_t)stream);
}

HRB_API int hrb_dense_bwd_x_t(const float* dz, int64_t lddz, const float* w, int64_t ldw, int64_t M, int32_t K, int32_t N,
                              const float* a_prev, int64_t lda_prev, int32_t act_prev, const uint32_t* relu_mask, float* dx, int64_t lddx,
                              float* dxt, int64_t lddxt, void* stream) {
  HRB_REQUIRE(dz && w && dx && M >= 0 && K > 0 && N > 0 && lddz >= N && ldw >= N && lddx >= K, "hrb_dense_bwd_x_t: bad argument");
  HRB_REQUIRE(dxt == nullptr || lddxt >= M, "hrb_dense_bwd_x_t: lddxt < M");
  if (M == 0) return HRB_OK;
  return hrb_tc_gemm_act_grad(dz, lddz, w, ldw, M, K, N, a_prev, lda_prev, act_prev, dx, lddx, dxt, lddxt, relu_mask, (K + 31) / 32,
                              (cudaStream_t)stream);
}

// y[M,N] = act(x[M,K] . wt[N,K]^T + group_bias[m / group_rows, :]): a Dense whose bias differs per group of consecutive rows
HRB_API int hrb_dense_fwd_t_grouped(const float* x, int64_t ldx, const float* wt, int64_t ldwt, const float* group_bias, int64_t ldgb,
                                    int32_t group_rows, int64_t M, int32_t K, int32_t N, int32_t act, float* y, int64_t ldy, void* stream) {
  HRB_REQUIRE(x && wt && y && group_bias && M >= 0 && K > 0 && N > 0 && ldx >= K && ldwt >= K && ldy >= N && ldgb >= N && group_rows > 0,
              "hrb_dense_fwd_t_grouped: bad argument");
  HRB_REQUIRE(act >= HRB_ACT_LINEAR && act <= HRB_ACT_TANH, "hrb_dense_fwd_t_grouped: unknown activation %d", act);
  if (M == 0) return HRB_OK;
  return hrb_tc_gemm_groupbias_act(x, ldx, wt, ldwt, group_bias, ldgb, group_rows, M, N, K, act, y, ldy, (cudaStream_t)stream);
}

HRB_API int hrb_dense_bwd_w_t_workspace(int64_t M, int32_t K, int32_t N, size_t* bytes) {
  HRB_REQUIRE(bytes && M >= 0 && K > 0 && N > 0, "hrb_dense_bwd_w_t_workspace: bad argument");
  const int splits = hrb_tc_splits(K, N, (int32_t)(M > 0x7fffffff ? 0x7fffffff : M));
  *bytes = (size_t)splits * K * (size_t)((N + 3) / 4 * 4) * sizeof(float) + (size_t)256 * N * sizeof(float) + 1024;
  return HRB_OK;
}

// dw[K,N] = xt[K,M] * dzt[N,M]^T (split over M, fixed-order reduce); dbias[N] = column sums of dz[M,N]
HRB_API int hrb_dense_bwd_w_t(const float* xt, int64_t ldxt, const float* dzt, int64_t lddzt, const float* dz, int64_t lddz, int64_t M,
                              int32_t K, int32_t N, float* dw, int64_t lddw, float* dbias, void* workspace, size_t workspace_bytes,
                              void* stream) {
  HRB_REQUIRE(xt && dzt && dw && workspace && M > 0 && K > 0 && N > 0 && ldxt >= M && lddzt >= M && lddw >= N, "hrb_dense_bwd_w_t: bad argument");
  (void)dz;
  (void)lddz;  // kept in the signature: the column sums now come from dz^T inside the GEMM
  HRB_REQUIRE(M <= 0x7fffffff, "hrb_dense_bwd_w_t: M too large");
  size_t need = 0;
  hrb_dense_bwd_w_t_workspace(M, K, N, &need);
  if (workspace_bytes < need) return fail(HRB_WORKSPACE, "hrb_dense_bwd_w_t: workspace %zu < required %zu bytes", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const int splits = hrb_tc_splits(K, N, (int32_t)M);
  const int64_t ldp = (N + 3) / 4 * 4;
  float* part = (float*)workspace;
  float* colpart = part + (size_t)splits * K * ldp;
  // the bias gradient (column sums of dz = row sums of dz^T) rides along in the GEMM: its converter warps touch every element
  // of the dz^T tiles anyway and leave one partial sum per (split, column)
  int rc = hrb_tc_gemm_splitk(xt, ldxt, dzt, lddzt, K, N, (int32_t)M, splits, part, ldp, dbias != nullptr ? colpart : nullptr, st);
  if (rc != HRB_OK) return rc;
  if (dbias != nullptr) {
    launch_split_reduce2(reduce_job(part, K, N, ldp, splits, dw, lddw), reduce_job(colpart, 1, N, N, splits, dbias, N), splits, st);
  } else {
    launch_split_reduce(part, K, N, ldp, splits, dw, lddw, st);
  }
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}

// The same weight gradient from x AS STORED (x[M,K] row-major, no x^T copy): dw[K,N] = x^T * dz with dz given as dzt[N,M]
HRB_API int hrb_dense_bwd_w_xn(const float* x, int64_t ldx, const float* dzt, int64_t lddzt, int64_t M, int32_t K, int32_t N, float* dw,
                               int64_t lddw, float* dbias, void* workspace, size_t workspace_bytes, void* stream) {
  HRB_REQUIRE(x && dzt && dw && workspace && M > 0 && K > 0 && N > 0 && ldx >= K && lddzt >= M && lddw >= N, "hrb_dense_bwd_w_xn: bad argument");
  HRB_REQUIRE(M <= 0x7fffffff, "hrb_dense_bwd_w_xn: M too large");
  size_t need = 0;
  hrb_dense_bwd_w_t_workspace(M, K, N, &need);
  if (workspace_bytes < need) return fail(HRB_WORKSPACE, "hrb_dense_bwd_w_xn: workspace %zu < required %zu bytes", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const int splits = hrb_tc_splits(K, N, (int32_t)M);
  const int64_t ldp = (N + 3) / 4 * 4;
  float* part = (float*)workspace;
  float* colpart = part + (size_t)splits * K * ldp;
  int rc = hrb_tc_gemm_splitk_an(x, ldx, dzt, lddzt, K, N, (int32_t)M, splits, part, ldp, dbias != nullptr ? colpart : nullptr, st);
  if (rc != HRB_OK) return rc;
  if (dbias != nullptr) {
    launch_split_reduce2(reduce_job(part, K, N, ldp, splits, dw, lddw), reduce_job(colpart, 1, N, N, splits, dbias, N), splits, st);
  } else {
    launch_split_reduce(part, K, N, ldp, splits, dw, lddw, st);
  }
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}

// ---------------------------------------------------------------------------------------------
// 1-unit Dense layer (the DeepFM logit layer): GEMV forward and a fused backward
// ---------------------------------------------------------------------------------------------
HRB_API int hrb_dense1_fwd(const float* x, int64_t ldx, const float* w, const float* bias, int64_t M, int32_t K, float* y, void* stream) {
  HRB_REQUIRE(M >= 0 && K > 0 && ldx >= K, "hrb_dense1_fwd: bad sizes");
  if (M == 0) return HRB_OK;
  HRB_REQUIRE(x && w && y, "hrb_dense1_fwd: null pointer");
  HRB_REQUIRE(aligned16(x) && aligned16(w) && ldx % 4 == 0, "hrb_dense1_fwd: x/w must be 16-byte aligned, ldx %% 4 == 0");
  const int64_t blocks = min((int64_t)sm_count() * 8, (M + 7) / 8);
  dense1_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, ldx, w, bias, M, K, y);
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}

HRB_API int hrb_dense1_bwd_workspace(int64_t M, int32_t K, size_t* bytes) {
  HRB_REQUIRE(bytes && M >= 0 && K > 0, "hrb_dense1_bwd_workspace: bad argument");
  *bytes = (size_t)512 * (K + 1) * sizeof(float) + 256;
  return HRB_OK;
}

// dz_prev[M,K] = dy[M] (x) w[K] * act'(x) (+ dz_prev^T), dw[K] = x^T dy, dbias = sum dy
HRB_API int hrb_dense1_bwd(const float* x, int64_t ldx, const float* w, const float* dy, int64_t M, int32_t K, int32_t act_prev,
                           float* dz_prev, int64_t lddz, float* dz_prev_t, int64_t lddzt, float* dw, float* dbias, void* workspace,
                           size_t workspace_bytes, void* stream) {
  HRB_REQUIRE(x && w && dy && dz_prev && dw && workspace && M > 0 && K > 0 && ldx >= K && lddz >= K, "hrb_dense1_bwd: bad argument");
  HRB_REQUIRE(dz_prev_t == nullptr || lddzt >= M, "hrb_dense1_bwd: lddzt < M");
  if (K % 4 != 0 || ldx % 4 != 0 || lddz % 4 != 0 || !aligned16(x) || !aligned16(w) || !aligned16(dz_prev))
    return fail(HRB_UNSUPPORTED, "hrb_dense1_bwd: needs K, ldx, lddz multiples of 4 and 16-byte aligned buffers");
  size_t need = 0;
  hrb_dense1_bwd_workspace(M, K, &need);
  if (workspace_bytes < need) return fail(HRB_WORKSPACE, "hrb_dense1_bwd: workspace %zu < required %zu bytes", workspace_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t fused_smem = ((size_t)2 * D1_ROWS * (K + 1) + D1_ROWS) * sizeof(float);
  if (K <= 190 && fused_smem <= 48 * 1024) {  // one pass: dz, dz^T and the dw / dbias partials of every CTA
    const int64_t tiles = (M + D1_ROWS - 1) / D1_ROWS;
    const int grid = (int)min((int64_t)min(512, sm_count() * 3), tiles);
    float* part = (float*)workspace;
    dense1_bwd_fused_kernel<<<grid, 256, fused_smem, st>>>(x, ldx, w, dy, M, K, act_prev, dz_prev, lddz, dz_prev_t, lddzt, part);
    HRB_LAUNCH_CHECK();
    if (dbias != nullptr) {
      launch_split_reduce2(reduce_job(part, 1, K, K + 1, grid, dw, K), reduce_job(part + K, 1, 1, K + 1, grid, dbias, 1), grid, st);
    } else {
      launch_split_reduce(part, 1, K, K + 1, grid, dw, K, st);
    }
    HRB_LAUNCH_CHECK();
    return HRB_OK;
  }
  const int64_t total = M * (K / 4);
  const int64_t blocks = min((int64_t)sm_count() * 8, (total + 255) / 256);
  dense1_bwd_dz_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, ldx, w, dy, M, K / 4, act_prev, dz_prev, lddz);
  HRB_LAUNCH_CHECK();
  if (dz_prev_t != nullptr) {
    dim3 tg((K + 31) / 32, (unsigned)((M + 31) / 32));
    if (tg.y > 65535) return fail(HRB_UNSUPPORTED, "hrb_dense1_bwd: M > 2M rows");
    hrb_transpose_kernel<<<tg, 256, 0, st>>>(dz_prev, M, K, lddz, dz_prev_t, lddzt);
    HRB_LAUNCH_CHECK();
  }
  int yb = (int)min((int64_t)512, (M + 127) / 128);
  if (yb < 1) yb = 1;
  const int64_t rpb = (M + yb - 1) / yb;
  float* part = (float*)workspace;
  dim3 grid((K + 31) / 32, yb);
  dense1_bwd_dw_kernel<<<grid, 256, 0, st>>>(x, ldx, dy, M, K, rpb, part);
  HRB_LAUNCH_CHECK();
  if (dbias != nullptr) {
    launch_split_reduce2(reduce_job(part, 1, K, K + 1, yb, dw, K), reduce_job(part + K, 1, 1, K + 1, yb, dbias, 1), yb, st);
  } else {
    launch_split_reduce(part, 1, K, K + 1, yb, dw, K, st);
  }
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}
