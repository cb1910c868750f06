// Register-tiled fp32 GEMM over shared-memory operands + tile loader, shared by the CUDA-core kernels.
#pragma once
#include "common.cuh"

namespace cfa {

constexpr int kNT = 256;             // threads per CTA (16 x 16 thread grid for the register-tiled GEMMs)

// ------------------------------------------------------------------------------------------------
// register-tiled GEMM over operands in shared memory.
// thread (ty,tx) = (tid/16, tid%16) owns rows m0+ty+16i (i<TM) and columns n0+tx+16j (j<TN).
// A(m,k) and Bm(k,n) are callables returning float; out-of-range rows/cols are clamped on read and
// dropped in tile_foreach.
// ------------------------------------------------------------------------------------------------
template <int TM, int TN, typename AF, typename BF>
__device__ __forceinline__ void tile_mac(float (&acc)[TM][TN], int m0, int n0, int M, int N, int K, AF A, BF Bm) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  int rm[TM], cn[TN];
#pragma unroll
  for (int i = 0; i < TM; ++i) rm[i] = min(m0 + ty + 16 * i, M - 1);
#pragma unroll
  for (int j = 0; j < TN; ++j) cn[j] = min(n0 + tx + 16 * j, N - 1);
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    float a[TM], b[TN];
#pragma unroll
    for (int i = 0; i < TM; ++i) a[i] = A(rm[i], k);
#pragma unroll
    for (int j = 0; j < TN; ++j) b[j] = Bm(k, cn[j]);
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

template <int TM, int TN, typename F>
__device__ __forceinline__ void tile_foreach(const float (&acc)[TM][TN], int m0, int n0, int M, int N, F f) {
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
      if (m < M && n < N) f(m, n, acc[i][j]);
    }
}

template <int TM, int TN>
__device__ __forceinline__ void tile_zero(float (&acc)[TM][TN]) {
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
}

// global [rows x D] (row stride D, element type T) columns [d0, d0+kc) -> smem fp32 [rows x ld], zero filled
// beyond D or beyond `rows_valid`.  128-bit (fp32) / 64-bit (16-bit types) loads when the layout allows.
template <typename T>
__device__ __forceinline__ void load_tile(float* __restrict__ dst, int ld, const T* __restrict__ src, int rows,
                                          int rows_valid, int D, int d0, int kc) {
  const bool vec = ((D & 3) == 0) && ((kc & 3) == 0) && ((((uintptr_t)src) & 15) == 0);
  if (vec) {
    const int q = kc >> 2;
    for (int idx = threadIdx.x; idx < rows * q; idx += kNT) {
      const int r = idx / q, c = (idx - r * q) << 2;
      float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
      if (r < rows_valid && d0 + c < D) {
        const T* p = src + (size_t)r * D + d0 + c;
        if constexpr (sizeof(T) == 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(p));
          x0 = t.x; x1 = t.y; x2 = t.z; x3 = t.w;
        } else {
          const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
          const T* h = reinterpret_cast<const T*>(&t);
          x0 = to_f32<T>(h[0]); x1 = to_f32<T>(h[1]); x2 = to_f32<T>(h[2]); x3 = to_f32<T>(h[3]);
        }
      }
      float* o = dst + r * ld + c;
      o[0] = x0; o[1] = x1; o[2] = x2; o[3] = x3;
    }
  } else {
    for (int idx = threadIdx.x; idx < rows * kc; idx += kNT) {
      const int r = idx / kc, c = idx - r * kc;
      float x = 0.f;
      if (r < rows_valid && d0 + c < D) x = to_f32<T>(src[(size_t)r * D + d0 + c]);
      dst[r * ld + c] = x;
    }
  }
}

}  // namespace cfa
