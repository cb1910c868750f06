// fp32-exact CUDA-core implementation of the SPARC + InfoNCE hot path (finetune/losses.py) for sm_100a.
//
// This is the full-precision path (fp32 FMA everywhere; parity target rtol 1e-5 vs the fp32 reference) and
// the fallback for shapes the tensor-core path does not cover.  One CTA per sample for the fine-grained
// part: the T x P similarity, its min-max normalisation, threshold, renormalised weights, the grouped patch
// embeddings and the T x T logits never leave shared memory / registers.  The backward is hand-derived
// (SURVEY.md §8 a-bwd) and works on the RAW embeddings: every gradient w.r.t. a normalised dot product is
// folded into a gradient w.r.t. the raw dot product, so all contractions read the raw input tiles.
#include "common.cuh"
#include "simt_tile.cuh"
#include "sparc_paths.h"
#include <math_constants.h>

namespace cfa {

constexpr float kNormEps = 1e-12f;   // F.normalize eps          (losses.py:152,173,207,212,221)
constexpr float kMinMaxEps = 1e-8f;  // losses.py:231
constexpr float kClampEps = 1e-8f;   // losses.py:211,242

// ------------------------------------------------------------------------------------------------
// SPARC fine-grained kernels
// ------------------------------------------------------------------------------------------------
struct SparcSmem {
  int ldS, ldT, db, ld;   // strides: S rows, logits rows (odd), D-block width, staging row stride (db+1)
  size_t S, dW, L, stV, stL, stG, stDG, vn, ln, gn2, msk, rs, cs, misc, total;   // float offsets
  float* gS;              // when the T x P tiles do not fit in shared memory (ViT-L/14@336: P = 577) they live in a
  float* gdW;             // per-CTA slice of a global scratch buffer (L2-resident); NULL = shared memory
};

// in_global: S (and dW) are kept in global scratch instead of shared memory
__host__ __device__ inline SparcSmem sparc_layout(int P, int T, int db, bool backward, bool in_global = false) {
  SparcSmem s;
  s.gS = nullptr; s.gdW = nullptr;
  s.ldS = P; s.ldT = T | 1; s.db = db; s.ld = db + 1;
  size_t o = 0;
  s.S = o; if (!in_global) o += (size_t)T * s.ldS;
  s.dW = o; if (backward && !in_global) o += (size_t)T * s.ldS;
  s.L = o; o += (size_t)T * s.ldT;
  s.stV = o; o += (size_t)P * s.ld;
  s.stL = o; o += (size_t)T * s.ld;
  s.stG = o; o += (size_t)T * s.ld;
  s.stDG = o; if (backward) o += (size_t)T * s.ld;
  s.vn = o; o += P;          // ||v_p|| (clamped)
  s.ln = o; o += T;          // ||l_t|| (clamped)
  s.gn2 = o; o += T;         // ||G_t||^2 accumulator, later ||G_t|| (clamped)
  s.msk = o; o += T;         // mask as float
  s.rs = o; o += (size_t)8 * T;   // per-row scalars: sigma, range, min, imin, imax, gfac, lfac, spare
  s.cs = o; o += P;          // per-column scalar (vfac)
  s.misc = o; o += 8 * 32 + 64;
  s.total = o;
  return s;
}

static int sparc_pick_db(int P, int T, bool backward, size_t limit_bytes, bool in_global = false) {
  for (int db = 32; db >= 8; db >>= 1)
    if (sparc_layout(P, T, db, backward, in_global).total * sizeof(float) <= limit_bytes) return db;
  return 0;
}

constexpr size_t kSmemLimit = 227 * 1024;

// Phase 1 (forward and backward): S_raw = l . v^T accumulated over D in chunks, row norms, optional pooled means.
template <typename T, bool kPool>
__device__ __forceinline__ void sparc_phase_similarity(const SparcSmem& L, float* sm, const T* vb, const T* lb, int P,
                                                       int Tn, int D, float* pooled_v, float* pooled_l, float cnt) {
  float* S = L.gS ? L.gS : sm + L.S; float* stV = sm + L.stV; float* stL = sm + L.stL;
  float* vn = sm + L.vn; float* ln = sm + L.ln; float* msk = sm + L.msk; float* red = sm + L.misc;
  const int ld = L.ld, kc = L.db;
  for (int i = threadIdx.x; i < P + Tn; i += kNT) { if (i < P) vn[i] = 0.f; else ln[i - P] = 0.f; }
  for (int d0 = 0; d0 < D; d0 += kc) {
    __syncthreads();
    load_tile<T>(stV, ld, vb, P, P, D, d0, kc);
    load_tile<T>(stL, ld, lb, Tn, Tn, D, d0, kc);
    __syncthreads();
    for (int r = threadIdx.x; r < P + Tn; r += kNT) {          // squared row norms (each row owned by one thread)
      const float* row = (r < P) ? stV + r * ld : stL + (r - P) * ld;
      float s = 0.f;
      for (int k = 0; k < kc; ++k) s = fmaf(row[k], row[k], s);
      if (r < P) vn[r] += s; else ln[r - P] += s;
    }
    if (kPool) {                                               // pooled means (losses.py:207,210-212)
      const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
      float sv = 0.f, sl = 0.f;
      if (c < kc) {
        for (int p = g; p < P; p += 8) sv += stV[p * ld + c];
        for (int t = g; t < Tn; t += 8) sl = fmaf(msk[t], stL[t * ld + c], sl);
      }
      red[g * 32 + c] = sv;
      __syncthreads();
      if (threadIdx.x < kc && d0 + threadIdx.x < D) {
        float s = 0.f;
        for (int k = 0; k < 8; ++k) s += red[k * 32 + threadIdx.x];
        pooled_v[d0 + threadIdx.x] = s / (float)P;
      }
      __syncthreads();
      red[g * 32 + c] = sl;
      __syncthreads();
      if (threadIdx.x < kc && d0 + threadIdx.x < D) {
        float s = 0.f;
        for (int k = 0; k < 8; ++k) s += red[k * 32 + threadIdx.x];
        pooled_l[d0 + threadIdx.x] = s / cnt;
      }
    }
    for (int m0 = 0; m0 < Tn; m0 += 80)
      for (int n0 = 0; n0 < P; n0 += 64) {
        float acc[5][4];
        tile_zero(acc);
        tile_mac<5, 4>(acc, m0, n0, Tn, P, kc, [&](int m, int k) { return stL[m * ld + k]; },
                       [&](int k, int n) { return stV[n * ld + k]; });
        if (d0 == 0) tile_foreach<5, 4>(acc, m0, n0, Tn, P, [&](int m, int n, float x) { S[m * L.ldS + n] = x; });
        else tile_foreach<5, 4>(acc, m0, n0, Tn, P, [&](int m, int n, float x) { S[m * L.ldS + n] += x; });
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P + Tn; i += kNT) {
    if (i < P) vn[i] = fmaxf(sqrtf(vn[i]), kNormEps); else ln[i - P] = fmaxf(sqrtf(ln[i - P]), kNormEps);
  }
  __syncthreads();
}

// Phase 2: row-wise min-max, threshold, renormalise (losses.py:228-243).  S_raw -> W in place.
// rs[t*8 + {0:sigma, 1:range, 2:min, 3:imin, 4:imax}] saved for the backward.
__device__ __forceinline__ void sparc_phase_weights(const SparcSmem& L, float* sm, int P, int Tn, float thr) {
  float* S = L.gS ? L.gS : sm + L.S; float* vn = sm + L.vn; float* ln = sm + L.ln; float* msk = sm + L.msk; float* rs = sm + L.rs;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int t = w; t < Tn; t += kNT / 32) {
    float* row = S + t * L.ldS;
    if (msk[t] == 0.f) {                      // masked token: contributes nothing ("truncate" semantics)
      for (int p = lane; p < P; p += 32) row[p] = 0.f;
      if (lane == 0) { rs[t * 8 + 0] = 1.f; rs[t * 8 + 1] = 1.f; rs[t * 8 + 2] = 0.f; rs[t * 8 + 3] = 0.f; rs[t * 8 + 4] = 0.f; }
      continue;
    }
    const float il = 1.f / ln[t];
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    int imn = 0, imx = 0;
    for (int p = lane; p < P; p += 32) {
      const float s = row[p] * il / vn[p];     // normalised similarity (losses.py:221-225)
      row[p] = s;
      if (s < mn) { mn = s; imn = p; }
      if (s > mx) { mx = s; imx = p; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {         // arg-min / arg-max, first occurrence on ties
      const float omn = __shfl_xor_sync(0xffffffffu, mn, o); const int oimn = __shfl_xor_sync(0xffffffffu, imn, o);
      const float omx = __shfl_xor_sync(0xffffffffu, mx, o); const int oimx = __shfl_xor_sync(0xffffffffu, imx, o);
      if (omn < mn || (omn == mn && oimn < imn)) { mn = omn; imn = oimn; }
      if (omx > mx || (omx == mx && oimx < imx)) { mx = omx; imx = oimx; }
    }
    const float rng = mx - mn + kMinMaxEps;    // losses.py:232
    float sum = 0.f;
    for (int p = lane; p < P; p += 32) {
      const float n = (row[p] - mn) / rng;
      const float th = (n < thr) ? 0.f : n;    // losses.py:235-239
      row[p] = th;
      sum += th;
    }
    sum = warp_sum(sum);
    const float sigma = fmaxf(sum, kClampEps); // losses.py:242
    for (int p = lane; p < P; p += 32) row[p] = row[p] / sigma;   // losses.py:243
    if (lane == 0) {
      rs[t * 8 + 0] = sigma; rs[t * 8 + 1] = rng; rs[t * 8 + 2] = mn;
      rs[t * 8 + 3] = __int_as_float(imn); rs[t * 8 + 4] = __int_as_float(imx);
    }
  }
  __syncthreads();
}

// G_blk = W . v_blk for the staged D-block (losses.py:245), rows of masked tokens zeroed -> stG.
__device__ __forceinline__ void sparc_group_block(const SparcSmem& L, float* sm, int P, int Tn) {
  float* W = L.gS ? L.gS : sm + L.S; float* stV = sm + L.stV; float* stG = sm + L.stG; float* msk = sm + L.msk;
  const int ld = L.ld, db = L.db;
  for (int m0 = 0; m0 < Tn; m0 += 80) {
    float acc[5][2];
    tile_zero(acc);
    tile_mac<5, 2>(acc, m0, 0, Tn, db, P, [&](int m, int k) { return W[m * L.ldS + k]; },
                   [&](int k, int n) { return stV[k * ld + n]; });
    tile_foreach<5, 2>(acc, m0, 0, Tn, db, [&](int m, int n, float x) { stG[m * ld + n] = x * msk[m]; });
  }
}

// Phase 3: over D-blocks: G_blk, ||G||^2, raw T x T logits  L_raw += G_blk . l_blk^T  (losses.py:173-180)
template <typename T>
__device__ __forceinline__ void sparc_phase_logits(const SparcSmem& L, float* sm, const T* vb, const T* lb, int P,
                                                   int Tn, int D) {
  float* Lg = sm + L.L; float* stV = sm + L.stV; float* stL = sm + L.stL; float* stG = sm + L.stG;
  float* gn2 = sm + L.gn2;
  const int ld = L.ld, db = L.db;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < Tn; i += kNT) gn2[i] = 0.f;
  for (int d0 = 0; d0 < D; d0 += db) {
    __syncthreads();
    load_tile<T>(stV, ld, vb, P, P, D, d0, db);
    load_tile<T>(stL, ld, lb, Tn, Tn, D, d0, db);
    __syncthreads();
    sparc_group_block(L, sm, P, Tn);
    __syncthreads();
    for (int t = w; t < Tn; t += kNT / 32) {
      float s = 0.f;
      for (int k = lane; k < db; k += 32) s = fmaf(stG[t * ld + k], stG[t * ld + k], s);
      s = warp_sum(s);
      if (lane == 0) gn2[t] += s;
    }
    for (int m0 = 0; m0 < Tn; m0 += 80)
      for (int n0 = 0; n0 < Tn; n0 += 64) {
        float acc[5][4];
        tile_zero(acc);
        tile_mac<5, 4>(acc, m0, n0, Tn, Tn, db, [&](int m, int k) { return stG[m * ld + k]; },
                       [&](int k, int n) { return stL[n * ld + k]; });
        if (d0 == 0) tile_foreach<5, 4>(acc, m0, n0, Tn, Tn, [&](int m, int n, float x) { Lg[m * L.ldT + n] = x; });
        else tile_foreach<5, 4>(acc, m0, n0, Tn, Tn, [&](int m, int n, float x) { Lg[m * L.ldT + n] += x; });
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Tn; i += kNT) gn2[i] = fmaxf(sqrtf(gn2[i]), kNormEps);   // now ||G_t||
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(kNT, 1)
sparc_fwd_kernel(const T* __restrict__ v, const T* __restrict__ l, const uint8_t* __restrict__ mask, int P, int Tn,
                 int D, int db, float thr, float scale, float* __restrict__ pooled_v, float* __restrict__ pooled_l,
                 float* __restrict__ lse_row, float* __restrict__ lse_col, float* __restrict__ local_partial,
                 float* scratch) {
  extern __shared__ float sm[];
  SparcSmem L = sparc_layout(P, Tn, db, false, scratch != nullptr);
  const int b = blockIdx.x;
  if (scratch) L.gS = scratch + (size_t)b * Tn * L.ldS;
  const T* vb = v + (size_t)b * P * D;
  const T* lb = l + (size_t)b * Tn * D;
  float* msk = sm + L.msk; float* ln = sm + L.ln; float* gn = sm + L.gn2; float* Lg = sm + L.L;
  float* red = sm + L.misc;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;

  for (int t = threadIdx.x; t < Tn; t += kNT) msk[t] = mask[(size_t)b * Tn + t] ? 1.f : 0.f;
  __syncthreads();
  float cnt = 0.f;
  for (int t = 0; t < Tn; ++t) cnt += msk[t];
  cnt = fmaxf(cnt, kClampEps);                                          // losses.py:211

  sparc_phase_similarity<T, true>(L, sm, vb, lb, P, Tn, D, pooled_v + (size_t)b * D, pooled_l + (size_t)b * D, cnt);
  sparc_phase_weights(L, sm, P, Tn, thr);
  sparc_phase_logits<T>(L, sm, vb, lb, P, Tn, D);

  // masked T x T logits, row / column log-sum-exp and CE (losses.py:177-196)
  float part_r = 0.f, part_c = 0.f;
  for (int i = w; i < Tn; i += kNT / 32) {
    if (msk[i] == 0.f) { if (lane == 0) lse_row[(size_t)b * Tn + i] = 0.f; continue; }
    float* row = Lg + i * L.ldT;
    float mx = -CUDART_INF_F;
    for (int j = lane; j < Tn; j += 32) {
      const float x = (msk[j] != 0.f) ? scale * (row[j] / (gn[i] * ln[j])) : -CUDART_INF_F;
      row[j] = x;
      mx = fmaxf(mx, x);
    }
    mx = warp_max(mx);
    float s = 0.f;
    for (int j = lane; j < Tn; j += 32) { const float x = row[j]; if (x != -CUDART_INF_F) s += expf(x - mx); }
    s = warp_sum(s);
    const float lse = mx + logf(s);
    __syncwarp();
    if (lane == 0) { lse_row[(size_t)b * Tn + i] = lse; part_r += lse - row[i]; }
  }
  __syncthreads();
  for (int j = w; j < Tn; j += kNT / 32) {
    if (msk[j] == 0.f) { if (lane == 0) lse_col[(size_t)b * Tn + j] = 0.f; continue; }
    float mx = -CUDART_INF_F;
    for (int i = lane; i < Tn; i += 32) if (msk[i] != 0.f) mx = fmaxf(mx, Lg[i * L.ldT + j]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int i = lane; i < Tn; i += 32) if (msk[i] != 0.f) s += expf(Lg[i * L.ldT + j] - mx);
    s = warp_sum(s);
    const float lse = mx + logf(s);
    if (lane == 0) { lse_col[(size_t)b * Tn + j] = lse; part_c += lse - Lg[j * L.ldT + j]; }
  }
  __syncthreads();
  if (lane == 0) { red[w] = part_r; red[8 + w] = part_c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int k = 0; k < kNT / 32; ++k) { a += red[k]; c += red[8 + k]; }
    local_partial[2 * b] = a;
    local_partial[2 * b + 1] = c;
  }
}

// dG_blk = dLhat . l_blk - G_blk * gfac  (masked rows zero) -> stDG.   Needs stG, stL staged and dLhat in L.
__device__ __forceinline__ void sparc_dgroup_block(const SparcSmem& L, float* sm, int Tn) {
  float* Lg = sm + L.L; float* stL = sm + L.stL; float* stG = sm + L.stG; float* stDG = sm + L.stDG;
  float* rs = sm + L.rs; float* msk = sm + L.msk;
  const int ld = L.ld, db = L.db;
  for (int m0 = 0; m0 < Tn; m0 += 80) {
    float acc[5][2];
    tile_zero(acc);
    tile_mac<5, 2>(acc, m0, 0, Tn, db, Tn, [&](int m, int k) { return Lg[m * L.ldT + k]; },
                   [&](int k, int n) { return stL[k * ld + n]; });
    tile_foreach<5, 2>(acc, m0, 0, Tn, db, [&](int m, int n, float x) {
      stDG[m * ld + n] = (x - stG[m * ld + n] * rs[m * 8 + 5]) * msk[m];
    });
  }
}

template <typename T>
__global__ void __launch_bounds__(kNT, 1)
sparc_bwd_kernel(const T* __restrict__ v, const T* __restrict__ l, const uint8_t* __restrict__ mask, int P, int Tn,
                 int D, int db, float thr, float scale, const float* __restrict__ lse_row,
                 const float* __restrict__ lse_col, const float* __restrict__ coef, const float* __restrict__ dpool_v,
                 const float* __restrict__ dpool_l, T* __restrict__ dv, T* __restrict__ dl, float* scratch) {
  extern __shared__ float sm[];
  SparcSmem L = sparc_layout(P, Tn, db, true, scratch != nullptr);
  const int b = blockIdx.x;
  if (scratch) { L.gS = scratch + (size_t)b * 2 * Tn * L.ldS; L.gdW = L.gS + (size_t)Tn * L.ldS; }
  const T* vb = v + (size_t)b * P * D;
  const T* lb = l + (size_t)b * Tn * D;
  T* dvb = dv + (size_t)b * P * D;
  T* dlb = dl + (size_t)b * Tn * D;
  float* W = L.gS ? L.gS : sm + L.S; float* dW = L.gdW ? L.gdW : sm + L.dW; float* Lg = sm + L.L;
  float* stV = sm + L.stV; float* stL = sm + L.stL; float* stG = sm + L.stG; float* stDG = sm + L.stDG;
  float* vn = sm + L.vn; float* ln = sm + L.ln; float* gn = sm + L.gn2; float* msk = sm + L.msk;
  float* rs = sm + L.rs; float* cs = sm + L.cs;
  const int ld = L.ld;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float c_r = coef[0], c_c = coef[1];

  for (int t = threadIdx.x; t < Tn; t += kNT) msk[t] = mask[(size_t)b * Tn + t] ? 1.f : 0.f;
  __syncthreads();
  float cnt = 0.f;
  for (int t = 0; t < Tn; ++t) cnt += msk[t];
  cnt = fmaxf(cnt, kClampEps);

  // ---- recompute the forward state
  sparc_phase_similarity<T, false>(L, sm, vb, lb, P, Tn, D, nullptr, nullptr, cnt);
  sparc_phase_weights(L, sm, P, Tn, thr);
  sparc_phase_logits<T>(L, sm, vb, lb, P, Tn, D);

  // ---- dL (SURVEY §8 a-bwd, masked local CE).  Column pass first (needs the raw logits): ldotL_j
  for (int j = w; j < Tn; j += kNT / 32) {
    float acc = 0.f;
    if (msk[j] != 0.f) {
      const float lc = lse_col[(size_t)b * Tn + j];
      for (int i = lane; i < Tn; i += 32) {
        if (msk[i] == 0.f) continue;
        const float x = scale * (Lg[i * L.ldT + j] / (gn[i] * ln[j]));
        float g = c_r * expf(x - lse_row[(size_t)b * Tn + i]) + c_c * expf(x - lc);
        if (i == j) g -= (c_r + c_c);
        acc = fmaf(g, x, acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) rs[j * 8 + 6] = acc;                 // ldotL_j = l^_j . dl^_j (logits part)
  }
  __syncthreads();
  // row pass: gdot_i, and dLhat (gradient w.r.t. the raw dot G_i . l_j) in place
  for (int i = w; i < Tn; i += kNT / 32) {
    float* row = Lg + i * L.ldT;
    float acc = 0.f;
    if (msk[i] == 0.f) {
      for (int j = lane; j < Tn; j += 32) row[j] = 0.f;
    } else {
      const float lr = lse_row[(size_t)b * Tn + i];
      for (int j = lane; j < Tn; j += 32) {
        float o = 0.f;
        if (msk[j] != 0.f) {
          const float den = gn[i] * ln[j];
          const float x = scale * (row[j] / den);
          float g = c_r * expf(x - lr) + c_c * expf(x - lse_col[(size_t)b * Tn + j]);
          if (i == j) g -= (c_r + c_c);
          acc = fmaf(g, x, acc);
          o = scale * g / den;
        }
        row[j] = o;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) rs[i * 8 + 5] = acc / (gn[i] * gn[i]);   // gfac_i = (g^_i . dg^_i) / ||G_i||^2
  }
  __syncthreads();

  // ---- dW = dG . v^T accumulated over D-blocks
  for (int d0 = 0; d0 < D; d0 += L.db) {
    __syncthreads();
    load_tile<T>(stV, ld, vb, P, P, D, d0, L.db);
    load_tile<T>(stL, ld, lb, Tn, Tn, D, d0, L.db);
    __syncthreads();
    sparc_group_block(L, sm, P, Tn);
    __syncthreads();
    sparc_dgroup_block(L, sm, Tn);
    __syncthreads();
    for (int m0 = 0; m0 < Tn; m0 += 80)
      for (int n0 = 0; n0 < P; n0 += 64) {
        float acc[5][4];
        tile_zero(acc);
        tile_mac<5, 4>(acc, m0, n0, Tn, P, L.db, [&](int m, int k) { return stDG[m * ld + k]; },
                       [&](int k, int n) { return stV[n * ld + k]; });
        if (d0 == 0) tile_foreach<5, 4>(acc, m0, n0, Tn, P, [&](int m, int n, float x) { dW[m * L.ldS + n] = x; });
        else tile_foreach<5, 4>(acc, m0, n0, Tn, P, [&](int m, int n, float x) { dW[m * L.ldS + n] += x; });
      }
  }
  __syncthreads();

  // ---- renorm / threshold / min-max backward, row pass: dW -> dShat (gradient w.r.t. raw l_t . v_p), ldotS_t
  for (int t = w; t < Tn; t += kNT / 32) {
    float* wr = W + t * L.ldS;
    float* gr = dW + t * L.ldS;
    if (msk[t] == 0.f) {
      for (int p = lane; p < P; p += 32) gr[p] = 0.f;
      if (lane == 0) rs[t * 8 + 6] = 0.f;
      continue;
    }
    const float sigma = rs[t * 8 + 0], rng = rs[t * 8 + 1], mn = rs[t * 8 + 2];
    const int imn = __float_as_int(rs[t * 8 + 3]), imx = __float_as_int(rs[t * 8 + 4]);
    float wdot = 0.f;
    for (int p = lane; p < P; p += 32) wdot = fmaf(wr[p], gr[p], wdot);
    wdot = warp_sum(wdot);
    float a1 = 0.f, a2 = 0.f;
    for (int p = lane; p < P; p += 32) {
      const bool kept = (thr <= 0.f) || (wr[p] > 0.f);
      const float dn = kept ? (gr[p] - wdot) / sigma : 0.f;
      const float n = wr[p] * sigma;
      a1 = fmaf(dn, n - 1.f, a1);
      a2 = fmaf(dn, n, a2);
      gr[p] = dn;
    }
    a1 = warp_sum(a1); a2 = warp_sum(a2);
    const float dmn = a1 / rng, dmx = -a2 / rng;
    float sdot = 0.f;
    for (int p = lane; p < P; p += 32) {
      const bool kept = (thr <= 0.f) || (wr[p] > 0.f);
      float ds = gr[p] / rng;
      if (p == imn) ds += dmn;
      if (p == imx) ds += dmx;
      const float s = kept ? fmaf(wr[p] * sigma, rng, mn) : mn;     // only read where ds != 0 (kept or argmin)
      sdot = fmaf(ds, s, sdot);
      gr[p] = ds / (ln[t] * vn[p]);
    }
    sdot = warp_sum(sdot);
    if (lane == 0) rs[t * 8 + 6] = (rs[t * 8 + 6] + sdot) / (ln[t] * ln[t]);   // lfac_t
  }
  __syncthreads();
  // column pass: vfac_p = (v^_p . dv^_p) / ||v_p||^2
  for (int p = threadIdx.x; p < P; p += kNT) {
    float acc = 0.f;
    for (int t = 0; t < Tn; ++t) {
      const float wv = W[t * L.ldS + p];
      const bool kept = (thr <= 0.f) || (wv > 0.f);
      const float s = kept ? fmaf(wv * rs[t * 8 + 0], rs[t * 8 + 1], rs[t * 8 + 2]) : rs[t * 8 + 2];
      acc = fmaf(dW[t * L.ldS + p] * (ln[t] * vn[p]), s, acc);
    }
    cs[p] = acc / (vn[p] * vn[p]);
  }
  __syncthreads();

  // ---- final pass over D-blocks: dv, dl
  const float invP = 1.f / (float)P, invc = 1.f / cnt;
  for (int d0 = 0; d0 < D; d0 += L.db) {
    __syncthreads();
    load_tile<T>(stV, ld, vb, P, P, D, d0, L.db);
    load_tile<T>(stL, ld, lb, Tn, Tn, D, d0, L.db);
    __syncthreads();
    sparc_group_block(L, sm, P, Tn);
    __syncthreads();
    sparc_dgroup_block(L, sm, Tn);
    __syncthreads();
    const int dn = min(L.db, D - d0);
    for (int m0 = 0; m0 < P; m0 += 80) {                 // dv block [P x db]
      float acc[5][2];
      tile_zero(acc);
      tile_mac<5, 2>(acc, m0, 0, P, L.db, Tn, [&](int m, int k) { return dW[k * L.ldS + m]; },
                     [&](int k, int n) { return stL[k * ld + n]; });
      tile_mac<5, 2>(acc, m0, 0, P, L.db, Tn, [&](int m, int k) { return W[k * L.ldS + m]; },
                     [&](int k, int n) { return stDG[k * ld + n]; });
      tile_foreach<5, 2>(acc, m0, 0, P, dn, [&](int m, int n, float x) {
        float o = x - stV[m * ld + n] * cs[m];
        if (dpool_v) o = fmaf(dpool_v[(size_t)b * D + d0 + n], invP, o);
        dvb[(size_t)m * D + d0 + n] = from_f32<T>(o);
      });
    }
    for (int m0 = 0; m0 < Tn; m0 += 80) {                // dl block [T x db]
      float acc[5][2];
      tile_zero(acc);
      tile_mac<5, 2>(acc, m0, 0, Tn, L.db, P, [&](int m, int k) { return dW[m * L.ldS + k]; },
                     [&](int k, int n) { return stV[k * ld + n]; });
      tile_mac<5, 2>(acc, m0, 0, Tn, L.db, Tn, [&](int m, int k) { return Lg[k * L.ldT + m]; },
                     [&](int k, int n) { return stG[k * ld + n]; });
      tile_foreach<5, 2>(acc, m0, 0, Tn, dn, [&](int m, int n, float x) {
        float o = x - stL[m * ld + n] * rs[m * 8 + 6];
        if (dpool_l) o = fmaf(dpool_l[(size_t)b * D + d0 + n] * msk[m], invc, o);
        dlb[(size_t)m * D + d0 + n] = from_f32<T>(o);
      });
    }
  }
}

// ------------------------------------------------------------------------------------------------
// scalar epilogues
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum_256(float x, float* red) {
  x = warp_sum(x);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
  __syncthreads();
  float s = 0.f;
  for (int k = 0; k < kNT / 32; ++k) s += red[k];
  return s;
}

__global__ void __launch_bounds__(kNT)
sparc_finalize_kernel(const float* global_sums, int global_batch, const float* local_partial, const uint8_t* mask,
                      int B, int Tn, float gw, float lw, float* out8, int gathered_ranks) {
  __shared__ float red[8];
  float nv = 0.f, a = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < B * Tn; i += kNT) nv += mask[i] ? 1.f : 0.f;
  for (int i = threadIdx.x; i < B; i += kNT) { a += local_partial[2 * i]; c += local_partial[2 * i + 1]; }
  nv = block_sum_256(nv, red);
  a = block_sum_256(a, red);
  c = block_sum_256(c, red);
  if (threadIdx.x == 0) {
    // mask.sum() + 1e-8 is evaluated in fp32 by the reference: the eps vanishes for counts >= 1 (losses.py:196)
    const float n_valid = nv + 1e-8f;
    float sa = global_sums[0], sb = global_sums[1];
    if (gathered_ranks > 1) {                                     // raw all-gather of per-rank [lse_a | lse_b | sum_a, sum_b]
      sa = 0.f; sb = 0.f;
      for (int r = 0; r < gathered_ranks; ++r) { sa += global_sums[(size_t)r * (2 * B + 2) + 2 * B]; sb += global_sums[(size_t)r * (2 * B + 2) + 2 * B + 1]; }
    }
    const float vl = sa / (float)global_batch;                    // losses.py:163
    const float lv = sb / (float)global_batch;
    const float vll = a / n_valid, lvl = c / n_valid;             // losses.py:196
    const float g = 0.5f * (vl + lv), lo = 0.5f * (vll + lvl);    // losses.py:217,252
    out8[0] = g; out8[1] = lo; out8[2] = gw * g + lw * lo;        // losses.py:254
    out8[3] = vl; out8[4] = lv; out8[5] = vll; out8[6] = lvl; out8[7] = n_valid;
  }
}

__global__ void sparc_coef_kernel(const float* grad7, float gw, float lw, int global_batch, const float* out8,
                                  float* coef8) {
  float u[7];
  for (int k = 0; k < 7; ++k) u[k] = grad7[k];
  const float gl = 0.5f * (u[0] + gw * u[2]);     // via global_loss and total_loss (losses.py:217,254)
  const float lo = 0.5f * (u[1] + lw * u[2]);     // via local_loss and total_loss  (losses.py:252,254)
  const float cvl = (u[3] + gl) / (float)global_batch, clv = (u[4] + gl) / (float)global_batch;
  coef8[0] = cvl; coef8[1] = clv;
  coef8[2] = (u[5] + lo) / out8[7];
  coef8[3] = (u[6] + lo) / out8[7];
  coef8[4] = clv; coef8[5] = cvl; coef8[6] = 0.f; coef8[7] = 0.f;
}

struct Grad7Ptrs { const float* g[7]; };

// same as sparc_coef_kernel, the 7 upstream gradients arriving as separate 0-dim tensors (NULL = not used)
__global__ void sparc_coef_ptrs_kernel(Grad7Ptrs gp, float gw, float lw, int global_batch, const float* out8, float* coef8,
                                       float gscale) {
  // gscale: world size when the caller's gradients are averaged over ranks afterwards (DDP), see cfa_sparc_loss_gathered_bwd_ex
  CoefSrc s{{gp.g[0], gp.g[1], gp.g[2], gp.g[3], gp.g[4], gp.g[5], gp.g[6]}, out8, gw, lw, gscale, global_batch, 1};
  float c[4];
  coef_from_src(s, c);
  coef8[0] = c[0]; coef8[1] = c[1];
  coef8[2] = c[2];
  coef8[3] = c[3];
  coef8[4] = c[1]; coef8[5] = c[0]; coef8[6] = 0.f; coef8[7] = 0.f;
}

}  // namespace cfa

using namespace cfa;

extern "C" int cfa_sparc_coef_ptrs(const float* g_global, const float* g_local, const float* g_total, const float* g_vl,
                                   const float* g_lv, const float* g_vl_local, const float* g_lv_local, float gw, float lw,
                                   int global_batch, const float* out8, float* coef8, void* stream) {
  if (!out8 || !coef8 || global_batch <= 0) return CFA_ERR_BAD_ARG;
  Grad7Ptrs gp{{g_global, g_local, g_total, g_vl, g_lv, g_vl_local, g_lv_local}};
  sparc_coef_ptrs_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(gp, gw, lw, global_batch, out8, coef8, 1.f);
  return launch_status();
}

namespace cfa {
int sparc_coef_ptrs_scaled(const float* const g[7], float gw, float lw, int global_batch, const float* out8, float* coef8,
                           float gscale, cudaStream_t st) {
  Grad7Ptrs gp{{g[0], g[1], g[2], g[3], g[4], g[5], g[6]}};
  sparc_coef_ptrs_kernel<<<1, 1, 0, st>>>(gp, gw, lw, global_batch, out8, coef8, gscale);
  return launch_status();
}
}  // namespace cfa

// bytes of global scratch the CUDA-core path needs for this shape (0 when the T x P tiles fit in shared memory)
extern "C" size_t cfa_sparc_scratch_bytes(int B, int P, int T, int backward) {
  if (B <= 0 || P <= 0 || T <= 0) return 0;
  if (sparc_pick_db(P, T, backward != 0, kSmemLimit) != 0) return 0;
  return (size_t)(backward ? 2 : 1) * B * T * P * sizeof(float);
}

extern "C" int cfa_sparc_max_patches(int T, int backward) {
  int best = 0;
  for (int P = 1; P <= 4096; ++P) {          // with the T x P tiles in global scratch only the staging tiles bound P
    if (sparc_pick_db(P, T, backward != 0, kSmemLimit) == 0 && sparc_pick_db(P, T, backward != 0, kSmemLimit, true) == 0) break;
    best = P;
  }
  return best;
}

template <typename T>
static int sparc_fwd_launch(const void* v, const void* l, const uint8_t* mask, int B, int P, int Tn, int D, float thr,
                            float scale, float* pooled_v, float* pooled_l, float* lse_row, float* lse_col,
                            float* local_partial, float* scratch, size_t scratch_bytes, cudaStream_t st) {
  int db = sparc_pick_db(P, Tn, false, kSmemLimit);
  bool in_global = false;
  if (db == 0) {                                   // T x P tile too large for shared memory: global scratch
    db = sparc_pick_db(P, Tn, false, kSmemLimit, true);
    if (db == 0) return CFA_ERR_UNSUPPORTED;
    if (!scratch || scratch_bytes < (size_t)B * Tn * P * sizeof(float)) return CFA_ERR_WORKSPACE;
    in_global = true;
  }
  const size_t smem = sparc_layout(P, Tn, db, false, in_global).total * sizeof(float);
  CFA_CUDA_TRY(cudaFuncSetAttribute(sparc_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sparc_fwd_kernel<T><<<B, kNT, smem, st>>>((const T*)v, (const T*)l, mask, P, Tn, D, db, thr, scale, pooled_v,
                                            pooled_l, lse_row, lse_col, local_partial, in_global ? scratch : nullptr);
  return launch_status();
}

int cfa::sparc_fwd_simt(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, int dtype,
                             float thr, float scale, float* pooled_v, float* pooled_l, float* lse_row, float* lse_col,
                             float* local_partial, void* scratch, size_t scratch_bytes, void* stream) {
  if (B <= 0 || P <= 0 || T <= 0 || D <= 0 || !v || !l || !mask) return CFA_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  float* sc = (float*)scratch;
  switch (dtype) {
    case CFA_DTYPE_F32: return sparc_fwd_launch<float>(v, l, mask, B, P, T, D, thr, scale, pooled_v, pooled_l, lse_row, lse_col, local_partial, sc, scratch_bytes, st);
    case CFA_DTYPE_BF16: return sparc_fwd_launch<__nv_bfloat16>(v, l, mask, B, P, T, D, thr, scale, pooled_v, pooled_l, lse_row, lse_col, local_partial, sc, scratch_bytes, st);
    case CFA_DTYPE_F16: return sparc_fwd_launch<__half>(v, l, mask, B, P, T, D, thr, scale, pooled_v, pooled_l, lse_row, lse_col, local_partial, sc, scratch_bytes, st);
    default: return CFA_ERR_UNSUPPORTED;
  }
}

template <typename T>
static int sparc_bwd_launch(const void* v, const void* l, const uint8_t* mask, int B, int P, int Tn, int D, float thr,
                            float scale, const float* lse_row, const float* lse_col, const float* coef,
                            const float* dpv, const float* dpl, void* dv, void* dl, float* scratch, size_t scratch_bytes,
                            cudaStream_t st) {
  int db = sparc_pick_db(P, Tn, true, kSmemLimit);
  bool in_global = false;
  if (db == 0) {
    db = sparc_pick_db(P, Tn, true, kSmemLimit, true);
    if (db == 0) return CFA_ERR_UNSUPPORTED;
    if (!scratch || scratch_bytes < (size_t)2 * B * Tn * P * sizeof(float)) return CFA_ERR_WORKSPACE;
    in_global = true;
  }
  const size_t smem = sparc_layout(P, Tn, db, true, in_global).total * sizeof(float);
  CFA_CUDA_TRY(cudaFuncSetAttribute(sparc_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sparc_bwd_kernel<T><<<B, kNT, smem, st>>>((const T*)v, (const T*)l, mask, P, Tn, D, db, thr, scale, lse_row, lse_col,
                                            coef, dpv, dpl, (T*)dv, (T*)dl, in_global ? scratch : nullptr);
  return launch_status();
}

int cfa::sparc_bwd_simt(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, int dtype,
                             float thr, float scale, const float* lse_row, const float* lse_col, const float* coef,
                             const float* dpooled_v, const float* dpooled_l, void* dv, void* dl, void* scratch,
                             size_t scratch_bytes, void* stream) {
  if (B <= 0 || P <= 0 || T <= 0 || D <= 0 || !v || !l || !mask || !coef || !dv || !dl) return CFA_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  float* sc = (float*)scratch;
  switch (dtype) {
    case CFA_DTYPE_F32: return sparc_bwd_launch<float>(v, l, mask, B, P, T, D, thr, scale, lse_row, lse_col, coef, dpooled_v, dpooled_l, dv, dl, sc, scratch_bytes, st);
    case CFA_DTYPE_BF16: return sparc_bwd_launch<__nv_bfloat16>(v, l, mask, B, P, T, D, thr, scale, lse_row, lse_col, coef, dpooled_v, dpooled_l, dv, dl, sc, scratch_bytes, st);
    case CFA_DTYPE_F16: return sparc_bwd_launch<__half>(v, l, mask, B, P, T, D, thr, scale, lse_row, lse_col, coef, dpooled_v, dpooled_l, dv, dl, sc, scratch_bytes, st);
    default: return CFA_ERR_UNSUPPORTED;
  }
}

extern "C" int cfa_sparc_finalize(const float* global_sums, int global_batch, const float* local_partial,
                                  const uint8_t* mask, int B, int T, float gw, float lw, float* out8, int gathered_ranks,
                                  void* stream) {
  if (B <= 0 || T <= 0 || global_batch <= 0) return CFA_ERR_BAD_ARG;
  sparc_finalize_kernel<<<1, kNT, 0, (cudaStream_t)stream>>>(global_sums, global_batch, local_partial, mask, B, T, gw,
                                                             lw, out8, gathered_ranks);
  return launch_status();
}

extern "C" int cfa_sparc_coef(const float* grad7, float gw, float lw, int global_batch, const float* out8,
                              float* coef8, void* stream) {
  if (!grad7 || !out8 || !coef8 || global_batch <= 0) return CFA_ERR_BAD_ARG;
  sparc_coef_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(grad7, gw, lw, global_batch, out8, coef8);
  return launch_status();
}
