// common.cu -- status strings, thread-local error detail, device attribute cache, table init.
#include <stdarg.h>

#include <atomic>

#include "common.cuh"

namespace hrb {

static thread_local char g_err[512] = "";

char* err_buf() { return g_err; }

int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return status;
}

static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

__global__ void init_uniform_kernel(float* __restrict__ t, int64_t rows, int32_t dim, uint32_t seed, float lo,
                                    float scale, int64_t row_start, int64_t row_step) {
  const int64_t n = rows * (int64_t)dim;
  const uint32_t salt = seed * 0x9E3779B9u + 0x85EBCA6Bu;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / dim;
    const int32_t c = (int32_t)(i - r * dim);
    const uint64_t idx = (uint64_t)(row_start + r * row_step) * (uint64_t)dim + (uint64_t)c;
    const uint32_t h = hash_u32((uint32_t)idx ^ hash_u32((uint32_t)(idx >> 32) + salt));
    const float u = __fmul_rn((float)(h >> 8), 5.9604644775390625e-08f);  // * 2^-24, exact
    t[i] = __fadd_rn(__fmul_rn(u, scale), lo);                             // two roundings, like numpy
  }
}

}  // namespace hrb

HRB_API int hrb_abi_version(void) { return HRB_ABI_VERSION; }

HRB_API const char* hrb_status_str(int status) {
  switch (status) {
    case HRB_OK: return "HRB_OK";
    case HRB_BAD_ARG: return "HRB_BAD_ARG";
    case HRB_UNSUPPORTED: return "HRB_UNSUPPORTED";
    case HRB_CUDA_ERROR: return "HRB_CUDA_ERROR";
    case HRB_WORKSPACE: return "HRB_WORKSPACE";
    default: return "HRB_UNKNOWN";
  }
}

HRB_API const char* hrb_last_error(void) { return hrb::err_buf(); }

namespace hrb { long long launches(); }
HRB_API int hrb_launch_count(int64_t* count) {
  HRB_REQUIRE(count != nullptr, "hrb_launch_count: count is NULL");
  *count = (int64_t)hrb::launches();
  return HRB_OK;
}

HRB_API int hrb_init_uniform(float* table, int64_t rows, int32_t dim, uint32_t seed, float lo, float hi,
                             int64_t row_start, int64_t row_step, void* stream) {
  HRB_REQUIRE(table != nullptr && rows >= 0 && dim > 0 && row_step > 0 && row_start >= 0,
              "hrb_init_uniform: bad argument");
  if (rows == 0) return HRB_OK;
  const int64_t n = rows * (int64_t)dim;
  const int threads = 256;
  int64_t blocks = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)hrb::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  hrb::init_uniform_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(table, rows, dim, seed, lo,
                                                                                  hi - lo, row_start, row_step);
  HRB_LAUNCH_CHECK();
  return HRB_OK;
}

// Let kernels launched on the current device dereference memory that lives on `peer_device` (NVLink / NVSwitch).
HRB_API int hrb_enable_peer_access(int32_t peer_device) {
  int cur = 0;
  HRB_CUDA(cudaGetDevice(&cur));
  if (cur == peer_device) return HRB_OK;
  int can = 0;
  HRB_CUDA(cudaDeviceCanAccessPeer(&can, cur, peer_device));
  if (!can) return hrb::fail(HRB_UNSUPPORTED, "device %d cannot access device %d peer-to-peer", cur, peer_device);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();  // clear the sticky-less error state
    return HRB_OK;
  }
  HRB_CUDA(e);
  return HRB_OK;
}
