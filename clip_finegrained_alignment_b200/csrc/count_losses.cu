// Counting losses of finetune/losses.py, fp32 CUDA-core kernels — small, latency-bound work:
//   count_contrastive : CountLoss's counterfactual term (losses.py:281-301): per sample, the image embedding against its
//                       caption (positive) and its C counterfactual captions; loss_b = log sum_c exp(e_i.e_cf[c]/T) -
//                       e_i.e_k/T on L2-normalised rows (the positive is NOT in the denominator, :295-298), mean over b.
//                       include_pos = 1 adds exp(pos) to the denominator: CLIPCountLoss.count_loss's grouping (:69-86).
//   logits_ce         : CountLoss's CLIP term on logits the CALLER computed (losses.py:276-279): mean row cross-entropy
//                       with diagonal targets of two [B,B] matrices, averaged.
#include "common.cuh"
#include <math_constants.h>

namespace cfa {

constexpr int kCcThreads = 128;

template <typename T>
__device__ __forceinline__ float cc_dot(const T* __restrict__ a, const T* __restrict__ b, int D, int lane) {
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s = fmaf(to_f32(a[d]), to_f32(b[d]), s);
  return warp_sum(s);
}

// one CTA per sample; scores[0] = positive, scores[1..C] = counterfactuals (raw dots / norms / T), norms kept in smem
template <typename T>
__device__ void cc_scores(const T* ei, const T* ek, const T* cf, int C, int D, float inv_t, float* sc, float* nrm) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { const float n = sqrtf(cc_dot(ei, ei, D, lane)); if (lane == 0) nrm[0] = n; }
  for (int c = warp; c < C + 1; c += kCcThreads / 32) {
    const T* row = c == 0 ? ek : cf + (size_t)(c - 1) * D;
    const float n = sqrtf(cc_dot(row, row, D, lane));
    const float dt = cc_dot(ei, row, D, lane);
    if (lane == 0) { nrm[1 + c] = n; sc[c] = dt; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C + 1; c += kCcThreads) sc[c] = sc[c] / (nrm[0] * nrm[1 + c]) * inv_t;
  __syncthreads();
}

__device__ __forceinline__ float cc_lse(const float* sc, int C, int include_pos) {   // every thread computes it (C is small)
  float m = -CUDART_INF_F;
  for (int c = include_pos ? 0 : 1; c < C + 1; ++c) m = fmaxf(m, sc[c]);
  float s = 0.f;
  for (int c = include_pos ? 0 : 1; c < C + 1; ++c) s += expf(sc[c] - m);
  return m + logf(s);
}

template <typename T>
__global__ void __launch_bounds__(kCcThreads)
count_contrastive_fwd_kernel(const T* __restrict__ ei, const T* __restrict__ ek, const T* __restrict__ cf, int C, int D,
                             float inv_t, int include_pos, float* __restrict__ per_sample) {
  extern __shared__ float sm[];
  float* sc = sm;                 // [C+1]
  float* nrm = sm + C + 1;        // [C+2]
  const int b = blockIdx.x;
  cc_scores(ei + (size_t)b * D, ek + (size_t)b * D, cf + (size_t)b * C * D, C, D, inv_t, sc, nrm);
  if (threadIdx.x == 0) per_sample[b] = cc_lse(sc, C, include_pos) - sc[0];
}

__global__ void __launch_bounds__(256)
mean_kernel(const float* __restrict__ x, int n, float scale, float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int k = 0; k < 8; ++k) t += red[k]; *out = t * scale; }
}

// backward: d pos = -g/B (+ g/B softmax_0 if include_pos), d cf_c = g/B softmax_c; through the dots and the normalisations
template <typename T>
__global__ void __launch_bounds__(kCcThreads)
count_contrastive_bwd_kernel(const T* __restrict__ ei, const T* __restrict__ ek, const T* __restrict__ cf, int B, int C,
                             int D, float inv_t, int include_pos, const float* __restrict__ grad, T* __restrict__ dei,
                             T* __restrict__ dek, T* __restrict__ dcf) {
  extern __shared__ float sm[];
  float* sc = sm;                 // [C+1] scores, then d score
  float* nrm = sm + C + 1;        // [C+2]
  float* acc = nrm + C + 2;       // [D]  d n_i accumulated in fp32
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* xi = ei + (size_t)b * D;
  const T* xk = ek + (size_t)b * D;
  const T* xc = cf + (size_t)b * C * D;
  cc_scores(xi, xk, xc, C, D, inv_t, sc, nrm);
  const float lse = cc_lse(sc, C, include_pos);
  const float gb = *grad / (float)B;
  __syncthreads();
  for (int c = threadIdx.x; c < C + 1; c += kCcThreads) {
    float d = (c >= 1 || include_pos) ? gb * expf(sc[c] - lse) : 0.f;
    if (c == 0) d -= gb;
    sc[c] = d * inv_t;            // d (n_i . n_c)
  }
  __syncthreads();
  const float ni = nrm[0];
  // rows: d n_c = dsc[c] * n_i  ->  J_n:  (d n_c - n_c (n_c . d n_c)) / |x_c| ;  n_c . d n_c = dsc[c] * (n_c . n_i)
  for (int c = warp; c < C + 1; c += kCcThreads / 32) {
    const T* row = c == 0 ? xk : xc + (size_t)(c - 1) * D;
    T* drow = c == 0 ? dek + (size_t)b * D : dcf + ((size_t)b * C + (c - 1)) * D;
    const float nc = nrm[1 + c], ds = sc[c];
    const float cosv = cc_dot(xi, row, D, lane) / (ni * nc);
    for (int d = lane; d < D; d += 32) {
      const float yi = to_f32(xi[d]) / ni, yc = to_f32(row[d]) / nc;
      drow[d] = from_f32<T>(ds * (yi - yc * cosv) / nc);
    }
  }
  // d n_i = sum_c dsc[c] * n_c, one thread per column in a fixed order (deterministic bits)
  for (int d = threadIdx.x; d < D; d += kCcThreads) {
    float a = 0.f;
    for (int c = 0; c < C + 1; ++c) {
      const T* row = c == 0 ? xk : xc + (size_t)(c - 1) * D;
      a = fmaf(sc[c], to_f32(row[d]) / nrm[1 + c], a);
    }
    acc[d] = a;
  }
  __syncthreads();
  // image row: J_n
  __shared__ float red[kCcThreads / 32];
  float dot = 0.f;
  for (int d = threadIdx.x; d < D; d += kCcThreads) dot = fmaf(acc[d], to_f32(xi[d]) / ni, dot);
  dot = warp_sum(dot);
  if (lane == 0) red[warp] = dot;
  __syncthreads();
  dot = red[0] + red[1] + red[2] + red[3];
  for (int d = threadIdx.x; d < D; d += kCcThreads)
    dei[(size_t)b * D + d] = from_f32<T>((acc[d] - to_f32(xi[d]) / ni * dot) / ni);
}

// ---- cross-entropy with diagonal targets on caller-provided logits: one warp per row of either matrix
template <typename T>
__global__ void __launch_bounds__(256)
logits_ce_fwd_kernel(const T* __restrict__ la, const T* __restrict__ lb, int B, float* __restrict__ lse2,
                     float* __restrict__ ce2 /* [2][B] */) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, which = blockIdx.y;
  if (r >= B) return;
  const T* row = (which ? lb : la) + (size_t)r * B;
  float m = -CUDART_INF_F;
  for (int j = lane; j < B; j += 32) m = fmaxf(m, to_f32(row[j]));
  m = warp_max(m);
  float s = 0.f;
  for (int j = lane; j < B; j += 32) s += expf(to_f32(row[j]) - m);
  s = warp_sum(s);
  if (lane == 0) {
    const float lse = m + logf(s);
    lse2[(size_t)which * B + r] = lse;
    ce2[(size_t)which * B + r] = lse - to_f32(row[r]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
logits_ce_bwd_kernel(const T* __restrict__ la, const T* __restrict__ lb, int B, const float* __restrict__ lse2,
                     const float* __restrict__ grad, T* __restrict__ dla, T* __restrict__ dlb) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, which = blockIdx.y;
  if (r >= B) return;
  const T* row = (which ? lb : la) + (size_t)r * B;
  T* drow = (which ? dlb : dla) + (size_t)r * B;
  const float lse = lse2[(size_t)which * B + r], c = *grad * 0.5f / (float)B;
  for (int j = lane; j < B; j += 32) {
    float g = expf(to_f32(row[j]) - lse);
    if (j == r) g -= 1.f;
    drow[j] = from_f32<T>(c * g);
  }
}

template <typename T>
static int cc_fwd(const void* ei, const void* ek, const void* cf, int B, int C, int D, float temperature, int include_pos,
                  float* per_sample, float* out, cudaStream_t st) {
  const size_t smem = sizeof(float) * (2 * C + 3);
  count_contrastive_fwd_kernel<T><<<B, kCcThreads, smem, st>>>((const T*)ei, (const T*)ek, (const T*)cf, C, D, 1.f / temperature,
                                                              include_pos, per_sample);
  mean_kernel<<<1, 256, 0, st>>>(per_sample, B, 1.f / (float)B, out);
  return launch_status();
}
template <typename T>
static int cc_bwd(const void* ei, const void* ek, const void* cf, int B, int C, int D, float temperature, int include_pos,
                  const float* grad, void* dei, void* dek, void* dcf, cudaStream_t st) {
  const size_t smem = sizeof(float) * (2 * C + 3 + D);
  if (smem > 48 * 1024) return CFA_ERR_UNSUPPORTED;
  count_contrastive_bwd_kernel<T><<<B, kCcThreads, smem, st>>>((const T*)ei, (const T*)ek, (const T*)cf, B, C, D,
                                                              1.f / temperature, include_pos, grad, (T*)dei, (T*)dek, (T*)dcf);
  return launch_status();
}

}  // namespace cfa

using namespace cfa;

extern "C" int cfa_count_contrastive_fwd(const void* ei, const void* ek, const void* ek_cf, int B, int C, int D, int dtype,
                                         float temperature, int include_pos, float* per_sample, float* out, void* stream) {
  if (B <= 0 || C <= 0 || D <= 0 || !ei || !ek || !ek_cf || !per_sample || !out || !(temperature > 0.f)) return CFA_ERR_BAD_ARG;
  if (C > 4096) return CFA_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case CFA_DTYPE_F32: return cc_fwd<float>(ei, ek, ek_cf, B, C, D, temperature, include_pos, per_sample, out, st);
    case CFA_DTYPE_BF16: return cc_fwd<__nv_bfloat16>(ei, ek, ek_cf, B, C, D, temperature, include_pos, per_sample, out, st);
    case CFA_DTYPE_F16: return cc_fwd<__half>(ei, ek, ek_cf, B, C, D, temperature, include_pos, per_sample, out, st);
    default: return CFA_ERR_UNSUPPORTED;
  }
}

extern "C" int cfa_count_contrastive_bwd(const void* ei, const void* ek, const void* ek_cf, int B, int C, int D, int dtype,
                                         float temperature, int include_pos, const float* grad, void* dei, void* dek,
                                         void* dek_cf, void* stream) {
  if (B <= 0 || C <= 0 || D <= 0 || !ei || !ek || !ek_cf || !grad || !dei || !dek || !dek_cf || !(temperature > 0.f))
    return CFA_ERR_BAD_ARG;
  if (C > 4096) return CFA_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case CFA_DTYPE_F32: return cc_bwd<float>(ei, ek, ek_cf, B, C, D, temperature, include_pos, grad, dei, dek, dek_cf, st);
    case CFA_DTYPE_BF16: return cc_bwd<__nv_bfloat16>(ei, ek, ek_cf, B, C, D, temperature, include_pos, grad, dei, dek, dek_cf, st);
    case CFA_DTYPE_F16: return cc_bwd<__half>(ei, ek, ek_cf, B, C, D, temperature, include_pos, grad, dei, dek, dek_cf, st);
    default: return CFA_ERR_UNSUPPORTED;
  }
}

extern "C" int cfa_logits_ce_fwd(const void* logits_a, const void* logits_b, int B, int dtype, float* lse2, float* ce2,
                                 float* out, void* stream) {
  if (B <= 0 || !logits_a || !logits_b || !lse2 || !ce2 || !out) return CFA_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid((B + 7) / 8, 2);
  switch (dtype) {
    case CFA_DTYPE_F32: logits_ce_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)logits_a, (const float*)logits_b, B, lse2, ce2); break;
    case CFA_DTYPE_BF16: logits_ce_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)logits_a, (const __nv_bfloat16*)logits_b, B, lse2, ce2); break;
    case CFA_DTYPE_F16: logits_ce_fwd_kernel<__half><<<grid, 256, 0, st>>>((const __half*)logits_a, (const __half*)logits_b, B, lse2, ce2); break;
    default: return CFA_ERR_UNSUPPORTED;
  }
  mean_kernel<<<1, 256, 0, st>>>(ce2, 2 * B, 0.5f / (float)B, out);      // (mean CE_a + mean CE_b) / 2
  return launch_status();
}

extern "C" int cfa_logits_ce_bwd(const void* logits_a, const void* logits_b, int B, int dtype, const float* lse2,
                                 const float* grad, void* dlogits_a, void* dlogits_b, void* stream) {
  if (B <= 0 || !logits_a || !logits_b || !lse2 || !grad || !dlogits_a || !dlogits_b) return CFA_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid((B + 7) / 8, 2);
  switch (dtype) {
    case CFA_DTYPE_F32: logits_ce_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)logits_a, (const float*)logits_b, B, lse2, grad, (float*)dlogits_a, (float*)dlogits_b); break;
    case CFA_DTYPE_BF16: logits_ce_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)logits_a, (const __nv_bfloat16*)logits_b, B, lse2, grad, (__nv_bfloat16*)dlogits_a, (__nv_bfloat16*)dlogits_b); break;
    case CFA_DTYPE_F16: logits_ce_bwd_kernel<__half><<<grid, 256, 0, st>>>((const __half*)logits_a, (const __half*)logits_b, B, lse2, grad, (__half*)dlogits_a, (__half*)dlogits_b); break;
    default: return CFA_ERR_UNSUPPORTED;
  }
  return launch_status();
}
