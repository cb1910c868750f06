// Shared device/host helpers for libcfa_b200 (sm_100a only).
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include "../../include/cfa_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libcfa_b200 is written for sm_100a (B200) only"
#endif

namespace cfa {

constexpr int kWarp = 32;

// Peer-memory gather (peer_exchange.cu): base[r] = rank r's exchange block mapped into this process (CUDA IPC over
// NVLink); n == 0: not used.  Passed by value to the kernels that read remote rows directly.
constexpr int kMaxPeers = 16;
struct PeerTable {
  const float* base[kMaxPeers];
  int n;
};

#define CFA_CUDA_TRY(expr)                         \
  do {                                             \
    cudaError_t _e = (expr);                       \
    if (_e != cudaSuccess) return (int)_e;         \
  } while (0)

// Upstream-gradient fan-in of the SPARC loss (autograd of losses.py:217,252,254-264) evaluated INSIDE the consuming kernels:
// the 7 upstream gradients arrive as device pointers (NULL = output unused), out8[7] is the number of valid tokens.
//   c[0], c[1]: coefficients of the two global InfoNCE directions (already divided by the global batch, times gscale)
//   c[2], c[3]: coefficients of the two token-level directions (already divided by n_valid)
// Same expressions as sparc_coef_ptrs_kernel (losses_simt.cu), which the per-stage entry points still launch.
struct CoefSrc {
  const float* g[7];
  const float* out8;
  float gw, lw, gscale;
  int global_batch;
  int on;                  // 0: the kernel reads a precomputed coefficient array instead
};
#ifdef __CUDACC__
__device__ __forceinline__ void coef_from_src(const CoefSrc& s, float c[4]) {
  float u[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) u[k] = s.g[k] ? *s.g[k] : 0.f;
  const float gl = 0.5f * (u[0] + s.gw * u[2]);
  const float lo = 0.5f * (u[1] + s.lw * u[2]);
  c[0] = s.gscale * (u[3] + gl) / (float)s.global_batch;
  c[1] = s.gscale * (u[4] + gl) / (float)s.global_batch;
  const float nv = s.out8[7];
  c[2] = (u[5] + lo) / nv;
  c[3] = (u[6] + lo) / nv;
}
#endif

// Programmatic dependent launch (griddepcontrol): a kernel launched with cfa_launch_pdl may START while its predecessor in the
// stream is still running -- once every CTA of the predecessor has executed pdl_launch_dependents() (or exited) -- and must
// call pdl_wait() before it touches anything the predecessor writes: the wait returns when the predecessor grid has
// completed and its memory is visible.  A predecessor that never triggers gives ordinary stream order; a kernel launched
// without the attribute sees both calls as no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// level: 1 = backward chain (the fine-grained backward under the global InfoNCE backward), 2 = forward chain (set-up of the
// short global InfoNCE kernels under their predecessors).  CFA_PDL = 0 / 1 / 2 enables levels <= its value; default 1.
// Measured at config 2 on one box (graph replay, ms per step): level 0: 0.2143, level 1: 0.2082, level 2: 0.2087 (the forward
// kernels need their predecessor's output at once, so only their few-hundred-cycle set-up overlaps: nothing gained).
static inline int cfa_pdl_level() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("CFA_PDL"); v = e ? atoi(e) : 1; }
  return v;
}
template <typename... KArgs, typename... Args>
static inline cudaError_t cfa_launch_pdl(int level, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                         Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = level <= cfa_pdl_level() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// Opt-in to > 48 KB of dynamic shared memory ONCE PER DEVICE (the attribute is per device: a per-process `static bool`
// guard would leave the second GPU of a multi-device process without it).  `mask` is the call site's static bit set.
#define CFA_SMEM_ATTR_ONCE(func, bytes)                                                                      \
  do {                                                                                                       \
    static unsigned long long cfa_attr_mask_ = 0ull;                                                         \
    int cfa_dev_ = 0;                                                                                        \
    CFA_CUDA_TRY(cudaGetDevice(&cfa_dev_));                                                                  \
    if (cfa_dev_ >= 64 || !((cfa_attr_mask_ >> cfa_dev_) & 1ull)) {                                          \
      CFA_CUDA_TRY(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));   \
      if (cfa_dev_ < 64) cfa_attr_mask_ |= 1ull << cfa_dev_;                                                 \
    }                                                                                                        \
  } while (0)


// launch check without synchronising
static inline int launch_status() { return (int)cudaGetLastError(); }

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}
__device__ __forceinline__ float warp_min(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fminf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}

template <typename T> __device__ __forceinline__ float to_f32(T x);
template <> __device__ __forceinline__ float to_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }
template <> __device__ __forceinline__ float to_f32<__half>(__half x) { return __half2float(x); }

template <typename T> __device__ __forceinline__ T from_f32(float x);
template <> __device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }
template <> __device__ __forceinline__ __half from_f32<__half>(float x) { return __float2half_rn(x); }

}  // namespace cfa
