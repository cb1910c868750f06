// SPARCLoss.masked_pairwise_contrastive_loss as a standalone entry point (finetune/losses.py:165-197):
// a, b [B,T,D], mask [B,T] -> sum_b sum_{valid i} CE_i( s * a^_i . b^_j over valid j ) / (sum mask + 1e-8).
// One direction (rows of a against columns of b).  fp32 CUDA-core tiles, one CTA per sample; the T x T logits stay in
// shared memory.  Inside SPARCLoss.forward the same computation is fused into cfa_sparc_fwd / cfa_sparc_bwd; this entry
// point exists so that the public helper of the reference keeps working on its own.  Masked tokens are skipped
// ("truncate" semantics): identical to the reference for all-True masks, finite where the reference is NaN.
#include "common.cuh"
#include "simt_tile.cuh"
#include <math_constants.h>

namespace cfa {

constexpr float kMpNormEps = 1e-12f;     // F.normalize eps (losses.py:173-174)
constexpr int kMpDb = 32;

struct MpSmem {
  int ldT, ld;
  size_t L, stA, stB, an, bn, msk, rdot, cdot, red, total;
};
__host__ __device__ inline MpSmem mp_layout(int T) {
  MpSmem s;
  s.ldT = T | 1; s.ld = kMpDb + 1;
  size_t o = 0;
  s.L = o; o += (size_t)T * s.ldT;
  s.stA = o; o += (size_t)T * s.ld;
  s.stB = o; o += (size_t)T * s.ld;
  s.an = o; o += T; s.bn = o; o += T; s.msk = o; o += T; s.rdot = o; o += T; s.cdot = o; o += T;
  s.red = o; o += 32;
  s.total = o;
  return s;
}

// raw logits L[i][j] = a_i . b_j and clamped norms
template <typename T>
__device__ __forceinline__ void mp_logits(const MpSmem& S, float* sm, const T* ab, const T* bb, int Tn, int D) {
  float* L = sm + S.L; float* stA = sm + S.stA; float* stB = sm + S.stB; float* an = sm + S.an; float* bn = sm + S.bn;
  for (int i = threadIdx.x; i < 2 * Tn; i += kNT) (i < Tn ? an[i] : bn[i - Tn]) = 0.f;
  for (int d0 = 0; d0 < D; d0 += kMpDb) {
    __syncthreads();
    load_tile<T>(stA, S.ld, ab, Tn, Tn, D, d0, kMpDb);
    load_tile<T>(stB, S.ld, bb, Tn, Tn, D, d0, kMpDb);
    __syncthreads();
    for (int r = threadIdx.x; r < 2 * Tn; r += kNT) {
      const float* row = (r < Tn) ? stA + r * S.ld : stB + (r - Tn) * S.ld;
      float s = 0.f;
      for (int k = 0; k < kMpDb; ++k) s = fmaf(row[k], row[k], s);
      if (r < Tn) an[r] += s; else bn[r - Tn] += s;
    }
    for (int m0 = 0; m0 < Tn; m0 += 80)
      for (int n0 = 0; n0 < Tn; n0 += 64) {
        float acc[5][4];
        tile_zero(acc);
        tile_mac<5, 4>(acc, m0, n0, Tn, Tn, kMpDb, [&](int m, int k) { return stA[m * S.ld + k]; },
                       [&](int k, int n) { return stB[n * S.ld + k]; });
        if (d0 == 0) tile_foreach<5, 4>(acc, m0, n0, Tn, Tn, [&](int m, int n, float x) { L[m * S.ldT + n] = x; });
        else tile_foreach<5, 4>(acc, m0, n0, Tn, Tn, [&](int m, int n, float x) { L[m * S.ldT + n] += x; });
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * Tn; i += kNT) {
    if (i < Tn) an[i] = fmaxf(sqrtf(an[i]), kMpNormEps); else bn[i - Tn] = fmaxf(sqrtf(bn[i - Tn]), kMpNormEps);
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(kNT, 1)
mp_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b, const uint8_t* __restrict__ mask, int Tn, int D, float scale,
              float* __restrict__ lse_row, float* __restrict__ partial) {
  extern __shared__ float sm[];
  const MpSmem S = mp_layout(Tn);
  const int bi = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* L = sm + S.L; float* an = sm + S.an; float* bn = sm + S.bn; float* msk = sm + S.msk; float* red = sm + S.red;
  for (int t = threadIdx.x; t < Tn; t += kNT) msk[t] = mask[(size_t)bi * Tn + t] ? 1.f : 0.f;
  mp_logits<T>(S, sm, a + (size_t)bi * Tn * D, b + (size_t)bi * Tn * D, Tn, D);
  float part = 0.f;
  for (int i = w; i < Tn; i += kNT / 32) {
    if (msk[i] == 0.f) { if (lane == 0) lse_row[(size_t)bi * Tn + i] = 0.f; continue; }
    float* row = L + i * S.ldT;
    float mx = -CUDART_INF_F;
    for (int j = lane; j < Tn; j += 32) {
      const float x = (msk[j] != 0.f) ? scale * (row[j] / (an[i] * bn[j])) : -CUDART_INF_F;   // losses.py:180,186
      row[j] = x;
      mx = fmaxf(mx, x);
    }
    mx = warp_max(mx);
    float s = 0.f;
    for (int j = lane; j < Tn; j += 32) { const float x = row[j]; if (x != -CUDART_INF_F) s += expf(x - mx); }
    s = warp_sum(s);
    const float lse = mx + logf(s);
    __syncwarp();
    if (lane == 0) { lse_row[(size_t)bi * Tn + i] = lse; part += lse - row[i]; }                   // losses.py:189-193
  }
  if (lane == 0) red[w] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < kNT / 32; ++k) s += red[k];
    partial[bi] = s;
  }
}

// out2[0] = sum partial / (sum mask + 1e-8), out2[1] = sum mask + 1e-8 (fp32, like the reference: losses.py:196)
__global__ void __launch_bounds__(kNT)
mp_finalize_kernel(const float* __restrict__ partial, const uint8_t* __restrict__ mask, int B, int Tn, float* __restrict__ out2) {
  __shared__ float red[2][8];
  float s = 0.f, nv = 0.f;
  for (int i = threadIdx.x; i < B; i += kNT) s += partial[i];
  for (int i = threadIdx.x; i < B * Tn; i += kNT) nv += mask[i] ? 1.f : 0.f;
  s = warp_sum(s); nv = warp_sum(nv);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = nv; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int k = 0; k < kNT / 32; ++k) { a += red[0][k]; c += red[1][k]; }
    const float n_valid = c + 1e-8f;
    out2[0] = a / n_valid;
    out2[1] = n_valid;
  }
}

template <typename T>
__global__ void __launch_bounds__(kNT, 1)
mp_bwd_kernel(const T* __restrict__ a, const T* __restrict__ b, const uint8_t* __restrict__ mask, int Tn, int D, float scale,
              const float* __restrict__ lse_row, const float* __restrict__ out2, const float* __restrict__ grad,
              T* __restrict__ da, T* __restrict__ db) {
  extern __shared__ float sm[];
  const MpSmem S = mp_layout(Tn);
  const int bi = blockIdx.x, lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* L = sm + S.L; float* stA = sm + S.stA; float* stB = sm + S.stB; float* an = sm + S.an; float* bn = sm + S.bn;
  float* msk = sm + S.msk; float* rdot = sm + S.rdot; float* cdot = sm + S.cdot;
  const T* ab = a + (size_t)bi * Tn * D;
  const T* bb = b + (size_t)bi * Tn * D;
  for (int t = threadIdx.x; t < Tn; t += kNT) msk[t] = mask[(size_t)bi * Tn + t] ? 1.f : 0.f;
  mp_logits<T>(S, sm, ab, bb, Tn, D);
  const float coef = grad[0] / out2[1];                 // d loss / d (sum of CE)
  // column pass first (needs the raw logits): cdot_j = sum_i dS_ij x_ij
  for (int j = w; j < Tn; j += kNT / 32) {
    float acc = 0.f;
    if (msk[j] != 0.f)
      for (int i = lane; i < Tn; i += 32) {
        if (msk[i] == 0.f) continue;
        const float x = scale * (L[i * S.ldT + j] / (an[i] * bn[j]));
        float g = coef * expf(x - lse_row[(size_t)bi * Tn + i]);
        if (i == j) g -= coef;
        acc = fmaf(g, x, acc);
      }
    acc = warp_sum(acc);
    if (lane == 0) cdot[j] = acc;
  }
  __syncthreads();
  // row pass: rdot_i and dLhat (gradient w.r.t. the raw dot a_i . b_j) in place
  for (int i = w; i < Tn; i += kNT / 32) {
    float* row = L + i * S.ldT;
    float acc = 0.f;
    const float lr = lse_row[(size_t)bi * Tn + i];
    for (int j = lane; j < Tn; j += 32) {
      float o = 0.f;
      if (msk[i] != 0.f && msk[j] != 0.f) {
        const float den = an[i] * bn[j];
        const float x = scale * (row[j] / den);
        float g = coef * expf(x - lr);
        if (i == j) g -= coef;
        acc = fmaf(g, x, acc);
        o = scale * g / den;
      }
      row[j] = o;
    }
    acc = warp_sum(acc);
    if (lane == 0) rdot[i] = acc;
  }
  __syncthreads();
  for (int d0 = 0; d0 < D; d0 += kMpDb) {
    __syncthreads();
    load_tile<T>(stA, S.ld, ab, Tn, Tn, D, d0, kMpDb);
    load_tile<T>(stB, S.ld, bb, Tn, Tn, D, d0, kMpDb);
    __syncthreads();
    const int dn = min(kMpDb, D - d0);
    for (int m0 = 0; m0 < Tn; m0 += 80) {
      float acc[5][2];
      tile_zero(acc);                                   // da block: sum_j dLhat_ij b_j - a_i rdot_i / |a_i|^2
      tile_mac<5, 2>(acc, m0, 0, Tn, kMpDb, Tn, [&](int m, int k) { return L[m * S.ldT + k]; },
                     [&](int k, int n) { return stB[k * S.ld + n]; });
      tile_foreach<5, 2>(acc, m0, 0, Tn, dn, [&](int m, int n, float x) {
        da[((size_t)bi * Tn + m) * D + d0 + n] = from_f32<T>(x - stA[m * S.ld + n] * (rdot[m] / (an[m] * an[m])));
      });
      tile_zero(acc);                                   // db block: sum_i dLhat_ij a_i - b_j cdot_j / |b_j|^2
      tile_mac<5, 2>(acc, m0, 0, Tn, kMpDb, Tn, [&](int m, int k) { return L[k * S.ldT + m]; },
                     [&](int k, int n) { return stA[k * S.ld + n]; });
      tile_foreach<5, 2>(acc, m0, 0, Tn, dn, [&](int m, int n, float x) {
        db[((size_t)bi * Tn + m) * D + d0 + n] = from_f32<T>(x - stB[m * S.ld + n] * (cdot[m] / (bn[m] * bn[m])));
      });
    }
  }
}

template <typename T>
static int mp_fwd_launch(const void* a, const void* b, const uint8_t* mask, int B, int Tn, int D, float scale, float* lse_row,
                         float* partial, float* out2, cudaStream_t st) {
  const size_t smem = mp_layout(Tn).total * sizeof(float);
  if (smem > 227 * 1024) return CFA_ERR_UNSUPPORTED;
  CFA_CUDA_TRY(cudaFuncSetAttribute(mp_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mp_fwd_kernel<T><<<B, kNT, smem, st>>>((const T*)a, (const T*)b, mask, Tn, D, scale, lse_row, partial);
  CFA_CUDA_TRY(cudaGetLastError());
  mp_finalize_kernel<<<1, kNT, 0, st>>>(partial, mask, B, Tn, out2);
  return launch_status();
}

template <typename T>
static int mp_bwd_launch(const void* a, const void* b, const uint8_t* mask, int B, int Tn, int D, float scale,
                         const float* lse_row, const float* out2, const float* grad, void* da, void* db, cudaStream_t st) {
  const size_t smem = mp_layout(Tn).total * sizeof(float);
  if (smem > 227 * 1024) return CFA_ERR_UNSUPPORTED;
  CFA_CUDA_TRY(cudaFuncSetAttribute(mp_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mp_bwd_kernel<T><<<B, kNT, smem, st>>>((const T*)a, (const T*)b, mask, Tn, D, scale, lse_row, out2, grad, (T*)da, (T*)db);
  return launch_status();
}

}  // namespace cfa

using namespace cfa;

extern "C" int cfa_masked_pairwise_fwd(const void* a, const void* b, const uint8_t* mask, int B, int T, int D, int dtype,
                                       float scale, float* lse_row, float* partial, float* out2, void* stream) {
  if (B <= 0 || T <= 0 || D <= 0 || !a || !b || !mask || !lse_row || !partial || !out2) return CFA_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case CFA_DTYPE_F32: return mp_fwd_launch<float>(a, b, mask, B, T, D, scale, lse_row, partial, out2, st);
    case CFA_DTYPE_BF16: return mp_fwd_launch<__nv_bfloat16>(a, b, mask, B, T, D, scale, lse_row, partial, out2, st);
    case CFA_DTYPE_F16: return mp_fwd_launch<__half>(a, b, mask, B, T, D, scale, lse_row, partial, out2, st);
    default: return CFA_ERR_UNSUPPORTED;
  }
}

extern "C" int cfa_masked_pairwise_bwd(const void* a, const void* b, const uint8_t* mask, int B, int T, int D, int dtype,
                                       float scale, const float* lse_row, const float* out2, const float* grad, void* da,
                                       void* db, void* stream) {
  if (B <= 0 || T <= 0 || D <= 0 || !a || !b || !mask || !lse_row || !out2 || !grad || !da || !db) return CFA_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case CFA_DTYPE_F32: return mp_bwd_launch<float>(a, b, mask, B, T, D, scale, lse_row, out2, grad, da, db, st);
    case CFA_DTYPE_BF16: return mp_bwd_launch<__nv_bfloat16>(a, b, mask, B, T, D, scale, lse_row, out2, grad, da, db, st);
    case CFA_DTYPE_F16: return mp_bwd_launch<__half>(a, b, mask, B, T, D, scale, lse_row, out2, grad, da, db, st);
    default: return CFA_ERR_UNSUPPORTED;
  }
}
