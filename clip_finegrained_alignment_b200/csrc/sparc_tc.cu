// Tensor-core (tcgen05 / TMEM / TMA) SPARC fine-grained kernels for bf16 embeddings on sm_100a.
//
// One CTA per sample, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner),
// warps 2..5 = epilogue (one thread per token row = one TMEM lane).  The raw bf16 l[T,D] and v[P,D] tiles
// stream through a TMA/mbarrier ring in 64-wide D blocks (SWIZZLE_128B) and are consumed directly by
// tcgen05.mma — normalisation is folded into the epilogue (S = (l.v)/(|l||v|)), so the bf16 products are
// exact and only fp32 accumulation order differs from the fp32 reference.  Operands produced on chip
// (alignment weights W, grouped embeddings G, gradient tiles) are split into bf16 hi + lo parts and fed as
// two MMAs, which keeps ~16 mantissa bits.  The T x P similarity never leaves TMEM / shared memory.
//
//   forward  pass 0:  S[T,P]  = sum_kb l_kb . v_kb^T                      (TMEM, 13x16 columns)
//            epilogue: min-max, threshold, renormalise -> W (hi/lo) in smem (losses.py:228-243)
//            pass 1:  G_kb    = W . v_kb       (v_kb tile read MN-major)  (losses.py:245)
//                     L[T,T] += G_kb . l_kb^T                             (losses.py:180)
//            epilogue: masked row/column log-sum-exp and CE               (losses.py:186-196)
#include "tc_common.cuh"
#include <cstdlib>
#include "sparc_paths.h"
#include <math_constants.h>

namespace cfa {
using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int kTcThreads = 192;
constexpr int kTcBwdThreads = 320;      // + 4 warps that only write dv / dl in the last pass
constexpr float kTcNormEps = 1e-12f, kTcMinMaxEps = 1e-8f, kTcClampEps = 1e-8f;
constexpr uint32_t kTmemCols = 512, kTmemG = 256, kTmemL = 384;

struct TcLayout {
  int NP, NT, KB, NS;
  uint32_t l_bytes, v_bytes, stage_bytes, w_bytes, g_bytes;
  uint32_t off_w, off_g, off_f, off_bar, total;     // byte offsets from the 1024-aligned base
};

__host__ __device__ inline TcLayout tc_fwd_layout(int P, int T, int D, int NS) {
  TcLayout L;
  L.NP = (P + 15) & ~15; L.NT = (T + 15) & ~15; L.KB = D / 64; L.NS = NS;
  L.l_bytes = L.NT * 128; L.v_bytes = L.NP * 128; L.stage_bytes = L.l_bytes + L.v_bytes;
  L.w_bytes = (uint32_t)L.NP * L.NT * 2;            // one of hi / lo, interleaved [NP/8][NT][8]
  L.g_bytes = 64u * L.NT * 2;                       // one of hi / lo of one G_kb buffer
  L.off_w = L.NS * L.stage_bytes;
  const uint32_t lb_bytes = (uint32_t)L.NT * (L.NT + 1) * 4;      // fp32 logits scratch aliases the W region
  L.off_g = L.off_w + ((2 * L.w_bytes > lb_bytes ? 2 * L.w_bytes : lb_bytes) + 1023 & ~1023u);
  L.off_f = L.off_g + 4 * L.g_bytes + 1024;         // +1 KB: phantom rows of the last interleaved chunk stay in bounds
  L.off_bar = L.off_f + 4 * ((L.NP + 32) + 3 * L.NT + 32);
  L.off_bar = (L.off_bar + 7) & ~7u;
  L.total = L.off_bar + 8 * (2 * L.NS + 12) + 16;
  return L;
}

// ------------------------------------------------------------------------------------------------
// prep: inverse row norms and pooled means in one pass over the inputs (losses.py:207-212, 221-222)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sparc_prep_kernel(const bf16* __restrict__ v, const bf16* __restrict__ l, const uint8_t* __restrict__ mask, int P, int T,
                  int D, float* __restrict__ inv_vn, float* __restrict__ inv_ln, float* __restrict__ pooled_v,
                  float* __restrict__ pooled_l) {
  extern __shared__ float red[];            // [8][D]
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ng = D >> 3;                    // 8-column groups per row (<= 128)
  for (int which = 0; which < 2; ++which) {
    const int rows = which ? T : P;
    const bf16* src = which ? l + (size_t)b * T * D : v + (size_t)b * P * D;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float cnt = 0.f;
    // 4 rows per warp iteration: all loads of the 4 rows are issued before any is consumed (memory-level parallelism)
    for (int r0 = warp * 4; r0 < rows; r0 += 32) {
      uint4 u[4][4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int g = lane + 32 * i, r = r0 + k;
          u[k][i] = (g < ng && r < rows) ? __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * D + 8 * g)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = r0 + k;
        const float m = (r < rows) ? (which ? (mask[(size_t)b * T + r] ? 1.f : 0.f) : 1.f) : 0.f;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bf16* h = reinterpret_cast<const bf16*>(&u[k][i]);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float f = __bfloat162float(h[j]);
            ss = fmaf(f, f, ss);
            acc[i][j] = fmaf(m, f, acc[i][j]);
          }
        }
        ss = warp_sum(ss);
        if (lane == 0 && r < rows) (which ? inv_ln + (size_t)b * T : inv_vn + (size_t)b * P)[r] = 1.f / fmaxf(sqrtf(ss), kTcNormEps);
      }
    }
    if (which) for (int t = 0; t < T; ++t) cnt += mask[(size_t)b * T + t] ? 1.f : 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int g = lane + 32 * i;
      if (g < ng)
#pragma unroll
        for (int j = 0; j < 8; ++j) red[warp * D + 8 * g + j] = acc[i][j];
    }
    __syncthreads();
    const float denom = which ? fmaxf(cnt, kTcClampEps) : (float)P;
    for (int d = threadIdx.x; d < D; d += 256) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += red[k * D + d];
      (which ? pooled_l : pooled_v)[(size_t)b * D + d] = s / denom;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// small helpers for the epilogue
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16x8(const float* x, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bf16 h0 = __float2bfloat16_rn(x[2 * i]), h1 = __float2bfloat16_rn(x[2 * i + 1]);
    const bf16 l0 = __float2bfloat16_rn(x[2 * i] - __bfloat162float(h0));
    const bf16 l1 = __float2bfloat16_rn(x[2 * i + 1] - __bfloat162float(h1));
    h[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    l[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// load `n` (16 or 32) consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, int n, float* x) {
  if (n == 32) tmem_ld32(taddr, x); else tmem_ld16(taddr, x);
  tmem_ld_wait();
}

struct TcFwdParams {
  int P, T, D, NS;
  float thr, scale;
  const uint8_t* mask;
  const float* inv_vn;
  const float* inv_ln;
  float* lse_row;
  float* lse_col;
  float* local_partial;
  float* tt_logits;     // [B][T][T] masked, scaled logits (saved for the backward), may be NULL
  float* g_inv_norm;    // [B][T]   1 / max(||G_t||, eps)
  bf16* g_split;        // [B][2][T][D] grouped embeddings G as bf16 hi | lo (saved for the backward), may be NULL
  float* q_save;        // [B][T][NP]   Q = G . v^T (raw), saved for the backward, may be NULL
};

template <int kNks>
__global__ void __launch_bounds__(kTcThreads, 1)
sparc_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmL, const TcFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const TcLayout L = tc_fwd_layout(p.P, p.T, p.D, p.NS);
  const int NP = L.NP, NT = L.NT, KB = L.KB, NS = L.NS, P = p.P, T = p.T;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* Whi = base + L.off_w;
  uint8_t* Wlo = Whi + L.w_bytes;
  uint8_t* Gs = base + L.off_g;                       // [buf][hi|lo][g_bytes]
  float* ivn = (float*)(base + L.off_f);              // [NP + 32], zero beyond P
  float* iln = ivn + NP + 32;                         // [NT]
  float* msk = iln + NT;                              // [NT]
  float* gnorm = msk + NT;                            // [NT]
  float* red = gnorm + NT;                            // [32]
  uint64_t* bars = (uint64_t*)(base + L.off_bar);
  uint64_t* full = bars;
  uint64_t* empty = bars + NS;
  uint64_t* s_full = bars + 2 * NS;
  uint64_t* w_ready = s_full + 1;
  uint64_t* g_full = s_full + 2;                      // [2]
  uint64_t* g_free = s_full + 4;                      // [2]
  uint64_t* gs_ready = s_full + 6;                    // [2]
  uint64_t* gs_free = s_full + 8;                     // [2]
  uint64_t* l_full = s_full + 10;
  uint32_t* tmem_slot = (uint32_t*)(s_full + 11);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(s_full, 1); mbar_init(w_ready, 4); mbar_init(l_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(g_full + i, 1); mbar_init(g_free + i, 4); mbar_init(gs_ready + i, 4); mbar_init(gs_free + i, 1); }
    fence_barrier_init();
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmL);
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  for (int i = threadIdx.x; i < NP + 32 + 2 * NT; i += kTcThreads) {
    if (i < NP + 32) ivn[i] = (i < P) ? p.inv_vn[(size_t)b * P + i] : 0.f;
    else if (i < NP + 32 + NT) { const int t = i - NP - 32; iln[t] = (t < T) ? p.inv_ln[(size_t)b * T + t] : 0.f; }
    else { const int t = i - NP - 32 - NT; msk[t] = (t < T && p.mask[(size_t)b * T + t]) ? 1.f : 0.f; }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      for (int u = 0; u < 2 * KB; ++u) {
        const int slot = u % NS, kb = u % KB;
        mbar_wait(empty + slot, ((u / NS) & 1) ^ 1);
        uint8_t* st = base + (size_t)slot * L.stage_bytes;
        mbar_expect_tx(full + slot, L.stage_bytes);
        tma_load_3d(st, &tmL, full + slot, kb * 64, 0, b);
        tma_load_3d(st + L.l_bytes, &tmV, full + slot, kb * 64, 0, b);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (warp-uniform control flow, one elected lane issues) ===============================
    const bool leader = elect_one();
    {
      const uint32_t idesc_s = make_idesc_bf16(128, NP, false, false);
      const uint32_t idesc_g = make_idesc_bf16(128, 64, false, true);
      const uint32_t idesc_l = make_idesc_bf16(128, NT, false, false);
      const uint32_t il_lbo = (uint32_t)NT * 16;      // interleaved operand: next 8-wide k chunk
      // Descriptors are built once; the issue loops only add to the 14-bit start-address field (units of 16 B),
      // so the single issuing thread spends a handful of instructions per MMA.
      const uint64_t sw0 = make_smem_desc(0, 16, 1024, kLayoutSw128);
      const uint64_t ilk_whi = make_smem_desc(smem_u32(Whi), il_lbo, 128, kLayoutNone);
      const uint64_t ilk_wlo = make_smem_desc(smem_u32(Wlo), il_lbo, 128, kLayoutNone);
      const uint32_t ilk_step = (2 * il_lbo) >> 4;
      // ---- pass 0: S = l . v^T
      for (int u = 0; u < KB; ++u) {
        const int slot = u % NS;
        mbar_wait(full + slot, (u / NS) & 1);
        tc_fence_after();
        const uint32_t sl = smem_u32(base + (size_t)slot * L.stage_bytes), sv = sl + L.l_bytes;
        const uint64_t dl0 = sw0 | (sl >> 4), dv0 = sw0 | (sv >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss_w(leader, tmem, dl0 + 2 * k, dv0 + 2 * k, idesc_s, (u | k) != 0);
        umma_commit_w(leader, empty + slot);
      }
      umma_commit_w(leader, s_full);
      // ---- pass 1: G_kb = W . v_kb ; L += G_kb . l_kb^T   (G issued one block ahead of L)
      mbar_wait(w_ready, 0);
      tc_fence_after();
      const int nks = kNks ? kNks : NP / 16;
      auto issue_g = [&](int kb) {
        const int u = KB + kb, slot = u % NS, buf = kb & 1;
        mbar_wait(full + slot, (u / NS) & 1);
        mbar_wait(g_free + buf, ((kb >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint64_t dv0 = sw0 | ((smem_u32(base + (size_t)slot * L.stage_bytes) + L.l_bytes) >> 4);
        const uint32_t d = tmem + kTmemG + 64 * buf;
        _Pragma("unroll") for (int ks = 0; ks < nks; ++ks) umma_ss_w(leader, d, ilk_whi + ks * ilk_step, dv0 + ks * 128, idesc_g, ks != 0);
        _Pragma("unroll") for (int ks = 0; ks < nks; ++ks) umma_ss_w(leader, d, ilk_wlo + ks * ilk_step, dv0 + ks * 128, idesc_g, true);
        umma_commit_w(leader, g_full + buf);
      };
      auto issue_l = [&](int kb) {
        const int u = KB + kb, slot = u % NS, buf = kb & 1;
        mbar_wait(gs_ready + buf, (kb >> 1) & 1);
        tc_fence_after();
        const uint64_t dl0 = sw0 | (smem_u32(base + (size_t)slot * L.stage_bytes) >> 4);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint64_t ga = make_smem_desc(smem_u32(Gs + (size_t)(2 * buf + half) * L.g_bytes), il_lbo, 128, kLayoutNone);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss_w(leader, tmem + kTmemL, ga + k * ilk_step, dl0 + 2 * k, idesc_l, (kb | half | k) != 0);
        }
        if (p.q_save) {       // Q += G_kb . v_kb^T into the (dead) S columns: the backward's dW = dLhat . S_raw - gfac * Q
          const uint64_t dv0 = sw0 | ((smem_u32(base + (size_t)slot * L.stage_bytes) + L.l_bytes) >> 4);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint64_t ga = make_smem_desc(smem_u32(Gs + (size_t)(2 * buf + half) * L.g_bytes), il_lbo, 128, kLayoutNone);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_ss_w(leader, tmem, ga + k * ilk_step, dv0 + 2 * k, idesc_s, (kb | half | k) != 0);
          }
        }
        umma_commit_w(leader, gs_free + buf);
        umma_commit_w(leader, empty + slot);
      };
      issue_g(0);
      for (int kb = 0; kb < KB; ++kb) {
        if (kb + 1 < KB) issue_g(kb + 1);
        issue_l(kb);
      }
      umma_commit_w(leader, l_full);
    }
  } else {
    // =============================== epilogue (4 warps, thread = token row) ===============================
    const int q = warp & 3;                           // TMEM lane quarter this warp may access
    const int row = 32 * q + lane;
    const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
    const bool valid = row < T && msk[row < NT ? row : 0] != 0.f;
    const float il = (row < NT) ? iln[row] : 0.f;

    // ---- epilogue 1: S -> W   (branch-free; columns >= P are masked with selects, never with control flow)
    mbar_wait(s_full, 0);
    tc_fence_after();
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    for (int c0 = 0; c0 < NP; c0 += 32) {
      float x[32];
      tmem_ld32(trow + c0, x);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float s = x[j] * il * ivn[c0 + j];
        const bool in = c0 + j < P;
        mn = fminf(mn, in ? s : CUDART_INF_F);
        mx = fmaxf(mx, in ? s : -CUDART_INF_F);
      }
    }
    const float inv_rng = 1.f / (mx - mn + kTcMinMaxEps);
    float sum = 0.f;
    for (int c0 = 0; c0 < NP; c0 += 32) {
      float x[32];
      tmem_ld32(trow + c0, x);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float nn = (x[j] * il * ivn[c0 + j] - mn) * inv_rng;
        sum += (c0 + j < P && !(nn < p.thr)) ? nn : 0.f;
      }
    }
    const float inv_sigma = valid ? 1.f / fmaxf(sum, kTcClampEps) : 0.f;
    for (int c0 = 0; c0 < NP; c0 += 32) {
      float x[32];
      tmem_ld32(trow + c0, x);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float nn = (x[j] * il * ivn[c0 + j] - mn) * inv_rng;
        x[j] = (valid && c0 + j < P && !(nn < p.thr)) ? nn * inv_sigma : 0.f;
      }
      if (row < NT) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (c0 + 8 * g < NP) {
            uint4 hi, lo;
            split_bf16x8(x + 8 * g, hi, lo);
            const uint32_t off = il_offset(NT, row, c0 + 8 * g);
            *reinterpret_cast<uint4*>(Whi + off) = hi;
            *reinterpret_cast<uint4*>(Wlo + off) = lo;
          }
        }
      }
    }
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(w_ready);

    // ---- epilogue 2: G_kb -> ||G||^2, bf16 hi/lo A operand for the logits MMA
    float gn2 = 0.f;
    for (int kb = 0; kb < KB; ++kb) {
      const int buf = kb & 1;
      mbar_wait(g_full + buf, (kb >> 1) & 1);
      tc_fence_after();
      float x[64];
      tmem_ld32(trow + kTmemG + 64 * buf, x);
      tmem_ld32(trow + kTmemG + 64 * buf + 32, x + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(g_free + buf);
      mbar_wait(gs_free + buf, ((kb >> 1) & 1) ^ 1);
      if (row < NT) {
        uint8_t* gh = Gs + (size_t)(2 * buf) * L.g_bytes;
        uint8_t* gl = gh + L.g_bytes;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float y[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { y[j] = valid ? x[8 * g + j] : 0.f; gn2 = fmaf(y[j], y[j], gn2); }
          uint4 hi, lo;
          split_bf16x8(y, hi, lo);
          const uint32_t off = il_offset(NT, row, 8 * g);
          *reinterpret_cast<uint4*>(gh + off) = hi;
          *reinterpret_cast<uint4*>(gl + off) = lo;
          if (p.g_split && row < T) {
            bf16* gdst = p.g_split + (((size_t)b * 2) * T + row) * p.D + kb * 64 + 8 * g;
            *reinterpret_cast<uint4*>(gdst) = hi;
            *reinterpret_cast<uint4*>(gdst + (size_t)T * p.D) = lo;
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(gs_ready + buf);
    }
    const float ign = 1.f / fmaxf(sqrtf(gn2), kTcNormEps);

    // ---- epilogue 3: masked logits, row LSE in registers, column LSE through smem (W region is free now)
    mbar_wait(l_full, 0);
    tc_fence_after();
    float* Lb = reinterpret_cast<float*>(Whi);          // [T][NT+1]
    const int ldl = NT + 1;
    float rmax = -CUDART_INF_F;
    const float sc_row = valid ? p.scale * ign : 0.f;
    for (int c0 = 0; c0 < NT; c0 += 16) {
      float x[16];
      tmem_ld_chunk(trow + kTmemL + c0, 16, x);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int col = c0 + j;
        const float y = (valid && msk[col] != 0.f) ? x[j] * sc_row * iln[col] : -CUDART_INF_F;   // msk/iln are 0 beyond T
        if (row < T) Lb[row * ldl + col] = y;
        rmax = fmaxf(rmax, y);
      }
    }
    if (p.g_inv_norm && row < T) p.g_inv_norm[(size_t)b * T + row] = ign;
    float ce_r = 0.f, ce_c = 0.f;
    if (valid) {
      float s = 0.f;
      for (int col = 0; col < T; ++col) { const float y = Lb[row * ldl + col]; if (y != -CUDART_INF_F) s += expf(y - rmax); }
      const float lse = rmax + logf(s);
      p.lse_row[(size_t)b * T + row] = lse;
      ce_r = lse - Lb[row * ldl + row];
    } else if (row < T) {
      p.lse_row[(size_t)b * T + row] = 0.f;
    }
    epi_bar_sync();
    if (valid) {                                         // thread `row` now owns column `row`
      float cmax = -CUDART_INF_F;
      for (int i = 0; i < T; ++i) cmax = fmaxf(cmax, Lb[i * ldl + row]);
      float s = 0.f;
      for (int i = 0; i < T; ++i) { const float y = Lb[i * ldl + row]; if (y != -CUDART_INF_F) s += expf(y - cmax); }
      const float lse = cmax + logf(s);
      p.lse_col[(size_t)b * T + row] = lse;
      ce_c = lse - Lb[row * ldl + row];
    } else if (row < T) {
      p.lse_col[(size_t)b * T + row] = 0.f;
    }
    ce_r = warp_sum(ce_r); ce_c = warp_sum(ce_c);
    if (lane == 0) { red[q] = ce_r; red[4 + q] = ce_c; }
    epi_bar_sync();
    if (row == 0) {
      p.local_partial[2 * b] = red[0] + red[1] + red[2] + red[3];
      p.local_partial[2 * b + 1] = red[4] + red[5] + red[6] + red[7];
    }
    if (p.q_save) {                                      // Q (TMEM columns of the dead S) -> global, row stride NP
      for (int c0 = 0; c0 < NP; c0 += 32) {
        float x[32];
        tmem_ld32(trow + c0, x);
        tmem_ld_wait();
        if (row < T) {
          float* qd = p.q_save + ((size_t)b * T + row) * NP + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (c0 + j < NP) *reinterpret_cast<float4*>(qd + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
        }
      }
    }
    if (p.tt_logits) {                                   // coalesced copy of the T x T logits for the backward
      float* dst = p.tt_logits + (size_t)b * T * T;
      for (int idx = threadIdx.x - 64; idx < T * T; idx += 128) { const int i = idx / T, j = idx - i * T; dst[idx] = Lb[i * ldl + j]; }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, kTmemCols); }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int tc_pick_stages(int P, int T, int D) {
  for (int ns = 4; ns >= 2; --ns)
    if (tc_fwd_layout(P, T, D, ns).total + 1024 <= 227 * 1024) return ns;
  return 0;
}

bool sparc_tc_supported(int P, int T, int D, int dtype) {
  if (dtype != CFA_DTYPE_BF16) return false;
  if (D % 256 != 0 || D > 1024 || T < 1 || T > 128 || P < 1 || P > 256) return false;
  return tc_pick_stages(P, T, D) != 0;
}

int sparc_prep_launch(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, float* inv_vn,
                      float* inv_ln, float* pooled_v, float* pooled_l, cudaStream_t st) {
  const size_t smem = (size_t)8 * D * sizeof(float);
  CFA_SMEM_ATTR_ONCE(sparc_prep_kernel, 8 * 1024 * 4);
  sparc_prep_kernel<<<B, 256, smem, st>>>((const bf16*)v, (const bf16*)l, mask, P, T, D, inv_vn, inv_ln, pooled_v, pooled_l);
  return launch_status();
}

int sparc_fwd_tc_launch(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, float thr,
                        float scale, float* row_inv_norm, float* pooled_v, float* pooled_l, float* lse_row,
                        float* lse_col, float* local_partial, float* tt_logits, float* g_inv_norm, void* g_split,
                        float* q_save, cudaStream_t st) {
  float* inv_vn = row_inv_norm;
  float* inv_ln = row_inv_norm + (size_t)B * P;
  int rc = sparc_prep_launch(v, l, mask, B, P, T, D, inv_vn, inv_ln, pooled_v, pooled_l, st);
  if (rc != CFA_OK) return rc;
  const int NS = tc_pick_stages(P, T, D);
  const TcLayout L = tc_fwd_layout(P, T, D, NS);
  CUtensorMap tmV, tmL;
  if ((rc = make_tmap_bf16_3d(&tmV, v, D, P, B, 64, L.NP)) != CFA_OK) return rc;
  if ((rc = make_tmap_bf16_3d(&tmL, l, D, T, B, 64, L.NT)) != CFA_OK) return rc;
  TcFwdParams prm{P, T, D, NS, thr, scale, mask, inv_vn, inv_ln, lse_row, lse_col, local_partial, tt_logits, g_inv_norm,
                  (bf16*)g_split, q_save};
  const size_t smem = L.total + 1024;
  if (L.NP == 208) {
    CFA_CUDA_TRY(cudaFuncSetAttribute(sparc_fwd_tc_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sparc_fwd_tc_kernel<13><<<B, kTcThreads, smem, st>>>(tmV, tmL, prm);
  } else {
    CFA_CUDA_TRY(cudaFuncSetAttribute(sparc_fwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sparc_fwd_tc_kernel<0><<<B, kTcThreads, smem, st>>>(tmV, tmL, prm);
  }
  return launch_status();
}


// ================================================================================================
// backward (tensor cores).  32-wide D blocks (SWIZZLE_64B TMA tiles) so that W, dShat (hi/lo), dLhat (hi/lo)
// and the per-block G / dG operands all stay resident in shared memory.  Four streaming passes over (l, v):
//   pass 1  S = l . v^T                                  -> epilogue: W (hi/lo), row stats
//   pass 2  G_kb = W . v_kb ; L += G_kb . l_kb^T         -> epilogue: dLhat (hi/lo), gfac, ldotL
//   pass 3  G_kb, X_kb = dLhat . l_kb -> dG_kb ; dW += dG_kb . v_kb^T   -> epilogue: dShat (hi/lo), lfac, vfac
//   pass 4  G_kb, dG_kb ; dv_kb = dShat^T . l_kb + W^T . dG_kb ; dl_kb = dShat . v_kb + dLhat^T . G_kb -> global
// All gradients are w.r.t. RAW dot products (SURVEY §8 a-bwd with the normalisations folded in), so every
// contraction reads the raw TMA tiles.
// ================================================================================================
struct TcBwdLayout {
  int NP, NT, KB, NS;
  uint32_t l_bytes, v_bytes, stage_bytes, w_bytes, dl_bytes, g_bytes;
  uint32_t off_w, off_dl, off_g, off_f, off_bar, total;
};

__host__ __device__ inline TcBwdLayout tc_bwd_layout(int P, int T, int D, int NS) {
  TcBwdLayout L;
  L.NP = (P + 15) & ~15; L.NT = (T + 15) & ~15; L.KB = D / 32; L.NS = NS;
  L.l_bytes = L.NT * 64; L.v_bytes = L.NP * 64; L.stage_bytes = L.l_bytes + L.v_bytes;
  L.w_bytes = (uint32_t)L.NP * L.NT * 2;
  L.dl_bytes = (uint32_t)L.NT * L.NT * 2;
  L.g_bytes = 32u * L.NT * 2;
  L.off_w = (L.NS * L.stage_bytes + 1023) & ~1023u;   // Whi, Wlo, DShi, DSlo
  const uint32_t sc_bytes = (uint32_t)L.NT * (L.NT + 1) * 4;       // phase-3 fp32 scratch aliases the dShat region
  L.off_dl = L.off_w + 2 * L.w_bytes + (2 * L.w_bytes > sc_bytes ? 2 * L.w_bytes : ((sc_bytes + 15) & ~15u));   // dLhi, dLlo
  L.off_g = L.off_dl + 2 * L.dl_bytes;                // 4 operand buffers of g_bytes
  L.off_f = L.off_g + 4 * L.g_bytes + 1024;           // phantom rows of the last buffer stay in bounds
  L.off_bar = (L.off_f + 4 * (2 * (L.NP + 32) + 7 * L.NT + 2 * D + 32) + 7) & ~7u;
  L.total = L.off_bar + 8 * (2 * L.NS + 34) + 16;
  return L;
}

// optional phase timestamps (clock64) for tuning: [B][2][16] long long, set through cfa_debug_set_profile_buffer
static long long* g_prof_buffer = nullptr;
static long long* g_prof_fwd = nullptr;     // same for the forward (cfa_debug_set_profile_buffer_fwd)

struct TcBwdParams {
  long long* prof;
  int P, T, D, NS;
  float thr, scale;
  const uint8_t* mask;
  const float* inv_vn;
  const float* inv_ln;
  const float* lse_row;
  const float* lse_col;
  const float* coef;
  const float* tt_logits;   // [B][T][T] from the forward
  const float* g_inv_norm;  // [B][T]
  const float* dpool_v;
  const float* dpool_l;
  const bf16* v;
  const bf16* l;
  bf16* dv;
  bf16* dl;
};

__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack_bf16x8(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    w[i] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(f[2 * i])) |
           ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(f[2 * i + 1])) << 16);
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// kNksP / kNksT: NP/16 and NT/16 as compile-time constants (fully unrolled issue loops with immediate descriptor
// increments), or 0 = runtime trip counts for other shapes.
template <int kNksP, int kNksT>
__global__ void __launch_bounds__(kTcBwdThreads, 1)
sparc_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmL, const TcBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const TcBwdLayout L = tc_bwd_layout(p.P, p.T, p.D, p.NS);
  const int NP = L.NP, NT = L.NT, KB = L.KB, NS = L.NS, P = p.P, T = p.T, D = p.D;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* Whi = base + L.off_w;
  uint8_t* Wlo = Whi + L.w_bytes;
  uint8_t* DShi = Wlo + L.w_bytes;
  uint8_t* DSlo = DShi + L.w_bytes;
  uint8_t* dLhi = base + L.off_dl;
  uint8_t* dLlo = dLhi + L.dl_bytes;
  uint8_t* Gb = base + L.off_g;                        // 4 x g_bytes
  float* ivn = (float*)(base + L.off_f);               // [NP + 32], zero beyond P
  float* vdot = ivn + NP + 32;                         // [NP + 32]  -> vfac
  float* iln = vdot + NP + 32;                         // [NT]
  float* msk = iln + NT;                               // [NT]
  float* lser = msk + NT;                              // [NT]
  float* lsec = lser + NT;                             // [NT]
  float* ldot = lsec + NT;                             // [NT]
  float* lfacs = ldot + NT;                            // [NT]   (l^_t . dl^_t) / |l_t|^2
  float* gfacs = lfacs + NT;                           // [NT]   (g^_t . dg^_t) / |G_t|^2
  float* dpool = gfacs + NT;                           // [2][D] gradient w.r.t. the pooled means of this sample
  uint64_t* bars = (uint64_t*)(base + L.off_bar);
  uint64_t* full = bars;
  uint64_t* empty = bars + NS;
  uint64_t* bb = bars + 2 * NS;
  uint64_t* s_full = bb + 0;  uint64_t* w_ready = bb + 1;
  uint64_t* g_full2 = bb + 2; uint64_t* g_free2 = bb + 4; uint64_t* gs_ready2 = bb + 6; uint64_t* gs_free2 = bb + 8;
  uint64_t* l_full = bb + 10; uint64_t* dl_ready = bb + 11;
  uint64_t* g_full3 = bb + 12; uint64_t* g_free3 = bb + 14; uint64_t* dg_ready3 = bb + 16; uint64_t* dg_free3 = bb + 18;
  uint64_t* dw_full = bb + 20; uint64_t* ds_ready = bb + 21;
  uint64_t* g_full4 = bb + 22; uint64_t* g_free4 = bb + 24; uint64_t* gs_ready4 = bb + 26; uint64_t* gs_free4 = bb + 28;
  uint64_t* out_full = bb + 30; uint64_t* out_free = bb + 32;
  uint32_t* tmem_slot = (uint32_t*)(bb + 36);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(s_full, 1); mbar_init(w_ready, 4); mbar_init(l_full, 1); mbar_init(dl_ready, 4);
    mbar_init(dw_full, 1); mbar_init(ds_ready, 4);
    for (int i = 0; i < 2; ++i) {
      mbar_init(gs_ready4 + i, 4); mbar_init(gs_free4 + i, 1);
      mbar_init(g_full2 + i, 1); mbar_init(g_free2 + i, 4); mbar_init(gs_ready2 + i, 4); mbar_init(gs_free2 + i, 1);
      mbar_init(g_full3 + i, 1); mbar_init(g_free3 + i, 4); mbar_init(dg_ready3 + i, 4); mbar_init(dg_free3 + i, 1);
      mbar_init(g_full4 + i, 1); mbar_init(g_free4 + i, 4); mbar_init(out_full + i, 1); mbar_init(out_free + i, 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmL);
  }
  if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
  for (int i = threadIdx.x; i < 2 * D; i += kTcBwdThreads) {
    const float* src = (i < D) ? p.dpool_v : p.dpool_l;
    dpool[i] = src ? src[(size_t)b * D + (i < D ? i : i - D)] : 0.f;
  }
  for (int i = threadIdx.x; i < 2 * (NP + 32) + 5 * NT; i += kTcBwdThreads) {
    if (i < NP + 32) ivn[i] = (i < P) ? p.inv_vn[(size_t)b * P + i] : 0.f;
    else if (i < 2 * (NP + 32)) vdot[i - NP - 32] = 0.f;
    else {
      const int k = (i - 2 * (NP + 32)) / NT, t = (i - 2 * (NP + 32)) % NT;
      const bool in = t < T;
      float x = 0.f;
      if (k == 0) x = in ? p.inv_ln[(size_t)b * T + t] : 0.f;
      else if (k == 1) x = (in && p.mask[(size_t)b * T + t]) ? 1.f : 0.f;
      else if (k == 2) x = in ? p.lse_row[(size_t)b * T + t] : 0.f;
      else if (k == 3) x = in ? p.lse_col[(size_t)b * T + t] : 0.f;
      (k == 0 ? iln : k == 1 ? msk : k == 2 ? lser : k == 3 ? lsec : ldot)[t] = x;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t cS = 0, cG = 256, cX = 320, cZ = 256, cDV = 0, cDL = 128;
  const uint32_t il_lbo = (uint32_t)NT * 16;           // interleaved operand: K-major chunk stride == MN-major group stride

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      for (int u = 0; u < 3 * KB; ++u) {
        const int slot = u % NS, kb = u % KB;
        mbar_wait(empty + slot, ((u / NS) & 1) ^ 1);
        uint8_t* st = base + (size_t)slot * L.stage_bytes;
        mbar_expect_tx(full + slot, L.stage_bytes);
        tma_load_3d(st, &tmL, full + slot, kb * 32, 0, b);
        tma_load_3d(st + L.l_bytes, &tmV, full + slot, kb * 32, 0, b);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (warp-uniform control flow, one elected lane issues) ===============================
    const bool leader = elect_one();
    {
      const uint32_t id_s = make_idesc_bf16(128, NP, false, false);     // [T x NP]  K-major x K-major
      const uint32_t id_kn32 = make_idesc_bf16(128, 32, false, true);   // A K-major, B MN-major, N = 32
      const uint32_t id_nn32 = make_idesc_bf16(128, 32, true, true);    // A MN-major, B MN-major, N = 32
      const int nksP = kNksP ? kNksP : NP / 16, nksT = kNksT ? kNksT : NT / 16;
      // Base descriptors built once; issue loops only add to the start-address field (units of 16 B).
      const uint64_t sw0 = make_smem_desc(0, 16, 512, kLayoutSw64);     // TMA tiles, K-major and MN-major alike
      auto ilk = [&](const uint8_t* a) { return make_smem_desc(smem_u32(a), il_lbo, 128, kLayoutNone); };   // interleaved, K-major
      auto ilm = [&](const uint8_t* a) { return make_smem_desc(smem_u32(a), 128, il_lbo, kLayoutNone); };   // interleaved, MN-major
      const uint32_t ks_k = (2 * il_lbo) >> 4;       // K-major interleaved: two 8-wide chunks per k-step
      const uint32_t ks_m = 256 >> 4;                // MN-major interleaved: 16 k-rows of 16 B
      const uint32_t mt_m = il_lbo;                  // MN-major interleaved: second 128-row M tile = 16 groups * il_lbo / 16
      const uint64_t k_whi = ilk(Whi), k_wlo = ilk(Wlo), k_dshi = ilk(DShi), k_dslo = ilk(DSlo);
      const uint64_t k_dlhi = ilk(dLhi), k_dllo = ilk(dLlo);
      const uint64_t m_whi = ilm(Whi), m_wlo = ilm(Wlo), m_dshi = ilm(DShi), m_dslo = ilm(DSlo);
      const uint64_t m_dlhi = ilm(dLhi), m_dllo = ilm(dLlo);
      const uint64_t k_g0 = ilk(Gb), k_g1 = ilk(Gb + L.g_bytes), k_g2 = ilk(Gb + 2 * L.g_bytes), k_g3 = ilk(Gb + 3 * L.g_bytes);
      const uint64_t m_g0 = ilm(Gb), m_g1 = ilm(Gb + L.g_bytes), m_g2 = ilm(Gb + 2 * L.g_bytes), m_g3 = ilm(Gb + 3 * L.g_bytes);
      long long* pf = (p.prof && leader) ? p.prof + (size_t)b * 32 : nullptr;
      int pi = 0;
      auto stamp = [&]() { if (pf) pf[pi++] = clock64(); };
      stamp();
      // ---- pass 1: S
      for (int u = 0; u < KB; ++u) {
        const int slot = u % NS;
        mbar_wait(full + slot, (u / NS) & 1);
        tc_fence_after();
        const uint32_t sl = smem_u32(base + (size_t)slot * L.stage_bytes), sv = sl + L.l_bytes;
        const uint64_t dl0 = sw0 | (sl >> 4), dv0 = sw0 | (sv >> 4);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_ss_w(leader, tmem + cS, dl0 + 2 * k, dv0 + 2 * k, id_s, (u | k) != 0);
        umma_commit_w(leader, empty + slot);
      }
      umma_commit_w(leader, s_full);
      stamp();

      // G_kb = W . v_kb (hi, lo) into TMEM cG[buf]; X_kb = dLhat . l_kb into cX[buf]
      auto issue_gx = [&](int pass, int kb, uint64_t* gfull, uint64_t* gfree) {
        const int u = pass * KB + kb, slot = u % NS, buf = kb & 1;
        mbar_wait(full + slot, (u / NS) & 1);
        mbar_wait(gfree + buf, ((kb >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t sl = smem_u32(base + (size_t)slot * L.stage_bytes), sv = sl + L.l_bytes;
        const uint64_t dl0 = sw0 | (sl >> 4), dv0 = sw0 | (sv >> 4);
        const uint32_t dg = tmem + cG + 32 * buf, dx = tmem + cX + 32 * buf;
        _Pragma("unroll") for (int ks = 0; ks < nksP; ++ks) umma_ss_w(leader, dg, k_whi + ks * ks_k, dv0 + ks * 64, id_kn32, ks != 0);
        _Pragma("unroll") for (int ks = 0; ks < nksP; ++ks) umma_ss_w(leader, dg, k_wlo + ks * ks_k, dv0 + ks * 64, id_kn32, true);
        _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, dx, k_dlhi + ks * ks_k, dl0 + ks * 64, id_kn32, ks != 0);
        _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, dx, k_dllo + ks * ks_k, dl0 + ks * 64, id_kn32, true);
        umma_commit_w(leader, gfull + buf);
      };

      // (the T x T logits and ||G|| come from the forward: no second pass)
      mbar_wait(w_ready, 0);
      tc_fence_after();
      stamp();
      stamp();
      // ---- pass 3: dW += dG_kb . v_kb^T
      mbar_wait(dl_ready, 0);
      tc_fence_after();
      stamp();
      issue_gx(1, 0, g_full3, g_free3);
      for (int kb = 0; kb < KB; ++kb) {
        if (kb + 1 < KB) issue_gx(1, kb + 1, g_full3, g_free3);
        const int u = KB + kb, slot = u % NS, buf = kb & 1;
        mbar_wait(dg_ready3 + buf, (kb >> 1) & 1);
        tc_fence_after();
        const uint64_t dv0 = sw0 | ((smem_u32(base + (size_t)slot * L.stage_bytes) + L.l_bytes) >> 4);
        const uint64_t a_hi = buf ? k_g2 : k_g0, a_lo = buf ? k_g3 : k_g1;
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_ss_w(leader, tmem + cS, a_hi + k * ks_k, dv0 + 2 * k, id_s, (kb | k) != 0);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_ss_w(leader, tmem + cS, a_lo + k * ks_k, dv0 + 2 * k, id_s, true);
        umma_commit_w(leader, dg_free3 + buf);
        umma_commit_w(leader, empty + slot);
      }
      // Z = dLhat^T . W  [T x NP]: folds  dl += dLhat^T . G  and the dLhat part of  dv += W^T . dG  into dShat' = dShat + Z
      {
        const uint32_t id_z = make_idesc_bf16(128, NP, true, true);
        _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, tmem + cZ, m_dlhi + ks * ks_m, m_whi + ks * ks_m, id_z, ks != 0);
        _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, tmem + cZ, m_dlhi + ks * ks_m, m_wlo + ks * ks_m, id_z, true);
        _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, tmem + cZ, m_dllo + ks * ks_m, m_whi + ks * ks_m, id_z, true);
      }
      umma_commit_w(leader, dw_full);
      stamp();

      // ---- pass 4: dv_kb = dShat'^T . l_kb + W^T . G'_kb ,  dl_kb = dShat' . v_kb      (G'_kb = -gfac * G_kb, hi/lo in Gb[2*buf..])
      mbar_wait(ds_ready, 0);
      tc_fence_after();
      stamp();
      auto issue_g4 = [&](int kb) {
        const int u = 2 * KB + kb, slot = u % NS, buf = kb & 1;
        mbar_wait(full + slot, (u / NS) & 1);
        mbar_wait(g_free4 + buf, ((kb >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint64_t dv0 = sw0 | ((smem_u32(base + (size_t)slot * L.stage_bytes) + L.l_bytes) >> 4);
        const uint32_t dg = tmem + cG + 32 * buf;
        _Pragma("unroll") for (int ks = 0; ks < nksP; ++ks) umma_ss_w(leader, dg, k_whi + ks * ks_k, dv0 + ks * 64, id_kn32, ks != 0);
        _Pragma("unroll") for (int ks = 0; ks < nksP; ++ks) umma_ss_w(leader, dg, k_wlo + ks * ks_k, dv0 + ks * 64, id_kn32, true);
        umma_commit_w(leader, g_full4 + buf);
      };
      issue_g4(0);
      for (int kb = 0; kb < KB; ++kb) {
        if (kb + 1 < KB) issue_g4(kb + 1);
        const int u = 2 * KB + kb, slot = u % NS, buf = kb & 1;
        mbar_wait(out_free + buf, ((kb >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t sl = smem_u32(base + (size_t)slot * L.stage_bytes), sv = sl + L.l_bytes;
        const uint64_t dl0 = sw0 | (sl >> 4), dv0 = sw0 | (sv >> 4);
        const uint64_t g_hi = buf ? m_g2 : m_g0, g_lo = buf ? m_g3 : m_g1;
        {                                             // dl rows t: issued first, it does not need G' (overlaps its conversion)
          const uint32_t d = tmem + cDL + 32 * buf;
          _Pragma("unroll") for (int ks = 0; ks < nksP; ++ks) umma_ss_w(leader, d, k_dshi + ks * ks_k, dv0 + ks * 64, id_kn32, ks != 0);
          _Pragma("unroll") for (int ks = 0; ks < nksP; ++ks) umma_ss_w(leader, d, k_dslo + ks * ks_k, dv0 + ks * 64, id_kn32, true);
        }
        mbar_wait(gs_ready4 + buf, (kb >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int m = 0; m < 2; ++m) {                 // dv rows p = 128 m ...
          const uint32_t d = tmem + cDV + 64 * buf + 32 * m;
          const uint32_t mo = m * mt_m;
          _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, d, m_dshi + mo + ks * ks_m, dl0 + ks * 64, id_nn32, ks != 0);
          _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, d, m_dslo + mo + ks * ks_m, dl0 + ks * 64, id_nn32, true);
          _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, d, m_whi + mo + ks * ks_m, g_hi + ks * ks_m, id_nn32, true);   // W^T . G': hi.hi
          _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, d, m_whi + mo + ks * ks_m, g_lo + ks * ks_m, id_nn32, true);   //           hi.lo
          _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, d, m_wlo + mo + ks * ks_m, g_hi + ks * ks_m, id_nn32, true);   //           lo.hi
        }
        umma_commit_w(leader, out_full + buf);
        umma_commit_w(leader, gs_free4 + buf);
        umma_commit_w(leader, empty + slot);
      }
      stamp();
    }
  } else if (warp < 6) {
    // =============================== epilogue A (4 warps, thread = TMEM lane) ===============================
    const int q = warp & 3;
    const int row = 32 * q + lane;
    const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
    const bool valid = row < T && msk[row < NT ? row : 0] != 0.f;
    const float il = (row < NT) ? iln[row] : 0.f;
    const float c_r = p.coef[0], c_c = p.coef[1];
    long long* pf = (p.prof && row == 0) ? p.prof + (size_t)b * 32 + 16 : nullptr;
    int pi = 0;
    auto stamp = [&]() { if (pf) pf[pi++] = clock64(); };
    stamp();
    // ---- phase 1: S -> W, row stats   (branch-free)
    mbar_wait(s_full, 0);
    tc_fence_after();
    stamp();
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    int imn = 0, imx = 0;
    for (int c0 = 0; c0 < NP; c0 += 32) {
      float x[32];
      tmem_ld32(trow + cS + c0, x);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float s = x[j] * il * ivn[c0 + j];
        const bool in = c0 + j < P;
        const bool lt = in && s < mn, gt = in && s > mx;     // strict: first occurrence wins, like torch.min/max
        mn = lt ? s : mn; imn = lt ? c0 + j : imn;
        mx = gt ? s : mx; imx = gt ? c0 + j : imx;
      }
    }
    const float rng = mx - mn + kTcMinMaxEps;
    const float inv_rng = 1.f / rng;
    float sum = 0.f;
    for (int c0 = 0; c0 < NP; c0 += 32) {
      float x[32];
      tmem_ld32(trow + cS + c0, x);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float nn = (x[j] * il * ivn[c0 + j] - mn) * inv_rng;
        sum += (c0 + j < P && !(nn < p.thr)) ? nn : 0.f;
      }
    }
    const float sigma = fmaxf(sum, kTcClampEps);
    const float inv_sigma = valid ? 1.f / sigma : 0.f;
    for (int c0 = 0; c0 < NP; c0 += 32) {
      float x[32];
      tmem_ld32(trow + cS + c0, x);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float nn = (x[j] * il * ivn[c0 + j] - mn) * inv_rng;
        x[j] = (valid && c0 + j < P && !(nn < p.thr)) ? nn * inv_sigma : 0.f;
      }
      if (row < NT) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (c0 + 8 * g < NP) {
            uint4 hi, lo;
            split_bf16x8(x + 8 * g, hi, lo);
            const uint32_t off = il_offset(NT, row, c0 + 8 * g);
            *reinterpret_cast<uint4*>(Whi + off) = hi;
            *reinterpret_cast<uint4*>(Wlo + off) = lo;
          }
      }
    }
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(w_ready);
    stamp();

    // ---- phase 3 (dLhat, gfac, ldotL) is done by the epilogue-B warps concurrently with pass 1 / phase 1
    stamp();
    stamp();
    asm volatile("bar.sync 3, 256;" ::: "memory");
    const float gfac = (row < NT) ? gfacs[row] : 0.f;
    const float ldl_j = (row < NT) ? ldot[row] : 0.f;
    stamp();

    // ---- pass 3 epilogue: dG_kb = (X_kb - G_kb gfac) -> A operand of the dW MMA
    for (int kb = 0; kb < KB; ++kb) {
      const int buf = kb & 1;
      mbar_wait(g_full3 + buf, (kb >> 1) & 1);
      tc_fence_after();
      float x[32], gx[32];
      tmem_ld32(trow + cG + 32 * buf, gx);
      tmem_ld32(trow + cX + 32 * buf, x);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(g_free3 + buf);
      mbar_wait(dg_free3 + buf, ((kb >> 1) & 1) ^ 1);
      if (row < NT) {
        uint8_t* gh = Gb + (size_t)(2 * buf) * L.g_bytes;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float y[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] = valid ? fmaf(-gx[8 * g + j], gfac, x[8 * g + j]) : 0.f;
          uint4 hi, lo;
          split_bf16x8(y, hi, lo);
          const uint32_t off = il_offset(NT, row, 8 * g);
          *reinterpret_cast<uint4*>(gh + off) = hi;
          *reinterpret_cast<uint4*>(gh + L.g_bytes + off) = lo;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(dg_ready3 + buf);
    }

    // ---- phase 5: dW -> dShat (hi/lo), sdot_t, vdot_p          (renorm / threshold / min-max backward)
    stamp();
    mbar_wait(dw_full, 0);
    tc_fence_after();
    stamp();
    auto load_w8 = [&](int c, float* w) {               // this row's W[c .. c+8) = hi + lo
      float h[8], l8[8];
      const uint32_t off = il_offset(NT, row < NT ? row : 0, c);
      unpack_bf16x8(*reinterpret_cast<const uint4*>(Whi + off), h);
      unpack_bf16x8(*reinterpret_cast<const uint4*>(Wlo + off), l8);
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = h[j] + l8[j];
    };
    float wdot = 0.f;
    for (int c0 = 0; c0 < NP; c0 += 32) {
      float x[32];
      tmem_ld32(trow + cS + c0, x);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (c0 + 8 * g < NP) {
          float w[8];
          load_w8(c0 + 8 * g, w);
#pragma unroll
          for (int j = 0; j < 8; ++j) wdot = fmaf(w[j], x[8 * g + j], wdot);     // W is 0 beyond P and on masked rows
        }
    }
    const float isg = 1.f / sigma;
    const bool keep_all = p.thr <= 0.f;
    float a1 = 0.f, a2 = 0.f;
    for (int c0 = 0; c0 < NP; c0 += 32) {
      float x[32];
      tmem_ld32(trow + cS + c0, x);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (c0 + 8 * g < NP) {
          float w[8];
          load_w8(c0 + 8 * g, w);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const bool kept = (keep_all || w[j] > 0.f) && (c0 + 8 * g + j < P);
            const float dn = kept ? (x[8 * g + j] - wdot) * isg : 0.f;
            const float nn = w[j] * sigma;
            a1 = fmaf(dn, nn - 1.f, a1);
            a2 = fmaf(dn, nn, a2);
          }
        }
    }
    const float dmn = a1 * inv_rng, dmx = -a2 * inv_rng;
    float sdot = 0.f;
    for (int c0 = 0; c0 < NP; c0 += 32) {
      float x[32], pr[32], z[32];
      tmem_ld32(trow + cS + c0, x);
      tmem_ld32(trow + cZ + c0, z);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (c0 + 8 * g < NP) {
          float w[8];
          load_w8(c0 + 8 * g, w);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int pc = c0 + 8 * g + j;
            const bool kept = (keep_all || w[j] > 0.f) && pc < P;
            float ds = kept ? ((x[8 * g + j] - wdot) * isg) * inv_rng : 0.f;
            ds += (pc == imn) ? dmn : 0.f;
            ds += (pc == imx) ? dmx : 0.f;
            ds = (valid && pc < P) ? ds : 0.f;
            const float sv = kept ? fmaf(w[j] * sigma, rng, mn) : mn;            // only read where ds != 0
            const float prod = (valid && pc < P) ? ds * sv : 0.f;   // phantom rows may hold NaN/Inf: select, never 0*x
            sdot += prod;
            pr[8 * g + j] = prod;
            x[8 * g + j] = (valid && pc < P) ? fmaf(ds * il, ivn[pc], z[8 * g + j]) : 0.f;   // dShat' = dShat + Z
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) pr[8 * g + j] = 0.f;
        }
      }
      if (row < NT) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (c0 + 8 * g < NP) {
            uint4 hi, lo;
            split_bf16x8(x + 8 * g, hi, lo);
            const uint32_t off = il_offset(NT, row, c0 + 8 * g);
            *reinterpret_cast<uint4*>(DShi + off) = hi;
            *reinterpret_cast<uint4*>(DSlo + off) = lo;
          }
      }
      // column sums over this warp's 32 rows, one shared-memory atomic per column per warp
      float mine = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float cs = warp_sum(pr[j]);
        mine = (lane == j) ? cs : mine;
      }
      if (c0 + lane < NP) atomicAdd(vdot + c0 + lane, mine);
    }
    if (row < NT) lfacs[row] = (sdot + ldl_j) * il * il;   // (l^_t . dl^_t) / ||l_t||^2
    epi_bar_sync();
    for (int i = threadIdx.x - 64; i < NP; i += 128) vdot[i] = vdot[i] * ivn[i] * ivn[i];   // -> vfac_p
    asm volatile("bar.sync 2, 256;" ::: "memory");      // vfac / lfac visible to the output warps (epilogue B)
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(ds_ready);
    stamp();

    // ---- pass 4, epilogue A: G'_kb = -gfac * G_kb (hi/lo) -> smem, double buffered (B operand of W^T . G')
    for (int kb = 0; kb < KB; ++kb) {
      const int buf = kb & 1;
      mbar_wait(g_full4 + buf, (kb >> 1) & 1);
      tc_fence_after();
      float gx[32];
      tmem_ld32(trow + cG + 32 * buf, gx);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(g_free4 + buf);
      mbar_wait(gs_free4 + buf, ((kb >> 1) & 1) ^ 1);
      if (row < NT) {
        uint8_t* gh = Gb + (size_t)(2 * buf) * L.g_bytes;
        const float ng = valid ? -gfac : 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float y[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] = valid ? gx[8 * g + j] * ng : 0.f;
          uint4 hi, lo;
          split_bf16x8(y, hi, lo);
          const uint32_t off = il_offset(NT, row, 8 * g);
          *reinterpret_cast<uint4*>(gh + off) = hi;
          *reinterpret_cast<uint4*>(gh + L.g_bytes + off) = lo;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(gs_ready4 + buf);
    }
    stamp();
    tc_fence_before();
  } else {
    // =============================== epilogue B (4 warps): phase 3, then dv_kb, dl_kb -> global ===============================
    const int q = warp & 3;
    const int row = 32 * q + lane;
    const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
    {
      // ---- phase 3: saved T x T logits -> dLhat (hi/lo), gfac_i, ldotL_j   (overlaps pass 1 and phase 1)
      const bool valid = row < T && msk[row < NT ? row : 0] != 0.f;
      const float c_r = p.coef[0], c_c = p.coef[1];
      float* Sc = reinterpret_cast<float*>(DShi);       // [T][NT+1] fp32 scratch, aliases the dShat region (free until phase 5)
      const int ldl = NT + 1;
      {
        const float* src = p.tt_logits + (size_t)b * T * T;
        for (int idx = threadIdx.x - 192; idx < T * T; idx += 128) { const int i = idx / T, j = idx - i * T; Sc[i * ldl + j] = __ldg(src + idx); }
      }
      const float ign = (row < T) ? p.g_inv_norm[(size_t)b * T + row] : 0.f;
      asm volatile("bar.sync 4, 128;" ::: "memory");
      float gdot = 0.f;
      const float lr = (row < NT) ? lser[row] : 0.f;
      for (int c0 = 0; c0 < NT; c0 += 16) {
        float x[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = c0 + j;
          const bool on = valid && col < T && msk[col] != 0.f;
          const float y = on ? Sc[row * ldl + col] : 0.f;
          const float den = ign * iln[col];
          float g = c_r * __expf(fminf(y - lr, 0.f)) + c_c * __expf(fminf(y - lsec[col], 0.f));
          g -= (col == row) ? (c_r + c_c) : 0.f;
          g = on ? g : 0.f;
          const float pr = g * y;
          gdot += pr;
          if (row < T && col < T) Sc[row * ldl + col] = pr;
          x[j] = p.scale * g * den;
        }
        if (row < NT) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            uint4 hi, lo;
            split_bf16x8(x + 8 * g, hi, lo);
            const uint32_t off = il_offset(NT, row, c0 + 8 * g);
            *reinterpret_cast<uint4*>(dLhi + off) = hi;
            *reinterpret_cast<uint4*>(dLlo + off) = lo;
          }
        }
      }
      if (row < NT) gfacs[row] = gdot * ign * ign;      // (g^_i . dg^_i) / ||G_i||^2
      asm volatile("bar.sync 4, 128;" ::: "memory");
      float ldl_j = 0.f;
      if (row < T) for (int i = 0; i < T; ++i) ldl_j += Sc[i * ldl + row];
      if (row < NT) ldot[row] = ldl_j;
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(dl_ready);
      asm volatile("bar.sync 3, 256;" ::: "memory");    // gfac / ldotL visible to epilogue A
    }
    asm volatile("bar.sync 2, 256;" ::: "memory");      // wait for vfac / lfac
    float cnt = 0.f;
    for (int t = 0; t < T; ++t) cnt += msk[t];
    const float invc = 1.f / fmaxf(cnt, kTcClampEps), invP = 1.f / (float)P;
    const float mrow = (row < NT) ? msk[row] * invc : 0.f;
    const float lfac = (row < NT) ? lfacs[row] : 0.f;
    const float vf0 = (row < P) ? vdot[row] : 0.f, vf1 = (128 + row < P) ? vdot[128 + row] : 0.f;
    for (int kb = 0; kb < KB; ++kb) {
      const int buf = kb & 1, d0 = kb * 32;
      mbar_wait(out_full + buf, (kb >> 1) & 1);
      tc_fence_after();
      float x0[32], x1[32], x2[32];
      tmem_ld32(trow + cDV + 64 * buf, x0);
      tmem_ld32(trow + cDV + 64 * buf + 32, x1);
      tmem_ld32(trow + cDL + 32 * buf, x2);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(out_free + buf);        // TMEM buffers are free as soon as they are in registers
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const int pr = 128 * m + row;
        if (pr < P) {
          const float* x = m ? x1 : x0;
          const float vf = m ? vf1 : vf0;
          const bf16* vsrc = p.v + ((size_t)b * P + pr) * D + d0;
          bf16* dst = p.dv + ((size_t)b * P + pr) * D + d0;
          uint4 raw[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) raw[g] = __ldg(reinterpret_cast<const uint4*>(vsrc) + g);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float vv[8], o[8];
            unpack_bf16x8(raw[g], vv);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaf(dpool[d0 + 8 * g + j], invP, fmaf(-vv[j], vf, x[8 * g + j]));
            reinterpret_cast<uint4*>(dst)[g] = pack_bf16x8(o);
          }
        }
      }
      if (row < T) {
        const bf16* lsrc = p.l + ((size_t)b * T + row) * D + d0;
        bf16* dst = p.dl + ((size_t)b * T + row) * D + d0;
        uint4 raw[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) raw[g] = __ldg(reinterpret_cast<const uint4*>(lsrc) + g);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float lv[8], o[8];
          unpack_bf16x8(raw[g], lv);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(dpool[D + d0 + 8 * g + j], mrow, fmaf(-lv[j], lfac, x2[8 * g + j]));
          reinterpret_cast<uint4*>(dst)[g] = pack_bf16x8(o);
        }
      }
    }
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, kTmemCols); }
}

static int tc_bwd_pick_stages(int P, int T, int D) {
  for (int ns = 4; ns >= 2; --ns)
    if (tc_bwd_layout(P, T, D, ns).total + 1024 <= 227 * 1024) return ns;
  return 0;
}

bool sparc_tc_bwd_supported(int P, int T, int D, int dtype) {
  if (!sparc_tc_supported(P, T, D, dtype)) return false;
  return tc_bwd_pick_stages(P, T, D) != 0;
}

int sparc_bwd_tc_launch(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, float thr,
                        float scale, const float* row_inv_norm, const float* lse_row, const float* lse_col,
                        const float* coef, const float* tt_logits, const float* g_inv_norm, const float* dpv,
                        const float* dpl, void* dv, void* dl, cudaStream_t st) {
  const int NS = tc_bwd_pick_stages(P, T, D);
  const TcBwdLayout L = tc_bwd_layout(P, T, D, NS);
  CUtensorMap tmV, tmL;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tmV, v, D, P, B, 32, L.NP)) != CFA_OK) return rc;
  if ((rc = make_tmap_bf16_3d(&tmL, l, D, T, B, 32, L.NT)) != CFA_OK) return rc;
  TcBwdParams prm{g_prof_buffer, P, T, D, NS, thr, scale, mask, row_inv_norm, row_inv_norm + (size_t)B * P, lse_row, lse_col, coef,
                  tt_logits, g_inv_norm, dpv, dpl, (const bf16*)v, (const bf16*)l, (bf16*)dv, (bf16*)dl};
  const size_t smem = L.total + 1024;
  if (L.NP == 208 && L.NT == 80) {          // ViT-B/16 (P = 196 / 197, T = 77): fully unrolled issue loops
    CFA_CUDA_TRY(cudaFuncSetAttribute(sparc_bwd_tc_kernel<13, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sparc_bwd_tc_kernel<13, 5><<<B, kTcBwdThreads, smem, st>>>(tmV, tmL, prm);
  } else {
    CFA_CUDA_TRY(cudaFuncSetAttribute(sparc_bwd_tc_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sparc_bwd_tc_kernel<0, 0><<<B, kTcBwdThreads, smem, st>>>(tmV, tmL, prm);
  }
  return launch_status();
}

}  // namespace cfa

namespace cfa {
thread_local bool g_sparc_bwd_pdl_late = false;
// generation switch (tuning aid): CFA_SPARC_GEN=2 keeps the second-generation kernels for A/B runs
bool sparc_gen3_enabled(int P, int T, int D, int dtype) {
  static int gen = -1;
  if (gen < 0) { const char* e = getenv("CFA_SPARC_GEN"); gen = (e && e[0] == '2') ? 2 : 3; }
  return gen == 3 && sparc_fwd3_supported(P, T, D, dtype) && sparc_bwd3_supported(P, T, D, dtype);
}
}  // namespace cfa

using namespace cfa;

// path: 0 = auto (tensor cores when the shape/dtype allows, else CUDA cores), 1 = CUDA cores, 2 = tensor cores
extern "C" int cfa_sparc_path(int P, int T, int D, int dtype, int path) {
  if (path == 1) return 1;
  const bool ok = sparc_tc_supported(P, T, D, dtype) || sparc_gen3_enabled(P, T, D, dtype);     // bf16; fp16 on the third generation only
  if (path == 2) return ok ? 2 : CFA_ERR_UNSUPPORTED;
  return ok ? 2 : 1;
}

// tuning aid: device buffer of [B][32] long long receiving clock64 phase stamps of the tensor-core backward
extern "C" int cfa_debug_set_profile_buffer(void* device_buffer) {
  g_prof_buffer = (long long*)device_buffer;
  return CFA_OK;
}

extern "C" int cfa_debug_set_profile_buffer_fwd(void* device_buffer) {
  g_prof_fwd = (long long*)device_buffer;
  return CFA_OK;
}

// same for the backward (the tensor-core backward needs more shared memory than the forward)
extern "C" int cfa_sparc_bwd_path(int P, int T, int D, int dtype, int path) {
  const int which = cfa_sparc_path(P, T, D, dtype, path);
  if (which != 2) return which;
  if (dtype == CFA_DTYPE_F16) return 2;
  return (sparc_gen3_enabled(P, T, D, dtype) || sparc_tc_bwd_supported(P, T, D, dtype) || sparc_bwd2_supported(P, T, D, dtype)) ? 2 : 1;
}

extern "C" int cfa_sparc_fwd(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, int dtype,
                             float thr, float scale, float* row_inv_norm, float* pooled_v, float* pooled_l,
                             float* lse_row, float* lse_col, float* local_partial, float* tt_logits, float* g_inv_norm,
                             void* g_split, float* q_save, void* scratch, size_t scratch_bytes, int path, void* stream) {
  if (B <= 0 || P <= 0 || T <= 0 || D <= 0 || !v || !l || !mask) return CFA_ERR_BAD_ARG;
  const int which = cfa_sparc_path(P, T, D, dtype, path);
  if (which < 0) return which;
  if (which == 2) {
    if (!row_inv_norm) return CFA_ERR_WORKSPACE;
    if ((g_split == nullptr) != (q_save == nullptr)) return CFA_ERR_BAD_ARG;
    if (g_split && sparc_gen3_enabled(P, T, D, dtype))   // third generation: transposed orientation, stacked hi|lo operands
      return sparc_fwd3_launch(v, l, mask, B, P, T, D, thr, scale, row_inv_norm, pooled_v, pooled_l, lse_row, lse_col,
                               local_partial, tt_logits, g_inv_norm, g_split, q_save, g_prof_fwd, dtype, (cudaStream_t)stream);
    if (!sparc_tc_supported(P, T, D, dtype)) return CFA_ERR_WORKSPACE;     // third-generation-only shape without its buffers
    if (sparc_fwd2_supported(P, T, D, dtype))          // one kernel: norms / pooled means fused into the streaming pass
      return sparc_fwd2_launch(v, l, mask, B, P, T, D, thr, scale, row_inv_norm, pooled_v, pooled_l, lse_row, lse_col,
                               local_partial, tt_logits, g_inv_norm, g_split, q_save, g_prof_fwd, dtype, (cudaStream_t)stream);
    return sparc_fwd_tc_launch(v, l, mask, B, P, T, D, thr, scale, row_inv_norm, pooled_v, pooled_l, lse_row, lse_col,
                               local_partial, tt_logits, g_inv_norm, g_split, q_save, (cudaStream_t)stream);
  }
  return sparc_fwd_simt(v, l, mask, B, P, T, D, dtype, thr, scale, pooled_v, pooled_l, lse_row, lse_col, local_partial,
                        scratch, scratch_bytes, stream);
}

extern "C" int cfa_sparc_bwd(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, int dtype,
                             float thr, float scale, const float* row_inv_norm, const float* lse_row,
                             const float* lse_col, const float* tt_logits, const float* g_inv_norm,
                             const void* g_split, const float* q_save, const float* coef, const float* dpooled_v,
                             const float* dpooled_l, void* dv, void* dl, void* scratch, size_t scratch_bytes, int path,
                             void* stream) {
  if (B <= 0 || P <= 0 || T <= 0 || D <= 0 || !v || !l || !mask || !coef || !dv || !dl) return CFA_ERR_BAD_ARG;
  const int which = cfa_sparc_bwd_path(P, T, D, dtype, path);
  if (which < 0) return which;
  if (which == 2) {
    if (!row_inv_norm || !tt_logits || !g_inv_norm) return CFA_ERR_WORKSPACE;
    if (dtype == CFA_DTYPE_F16 && (!g_split || !q_save)) return CFA_ERR_WORKSPACE;      // fp16: third generation only (needs its buffers)
    if (g_split && q_save && sparc_gen3_enabled(P, T, D, dtype))
      return sparc_bwd3_launch(v, l, mask, B, P, T, D, thr, scale, row_inv_norm, lse_row, lse_col, coef, tt_logits,
                               g_inv_norm, g_split, q_save, dpooled_v, dpooled_l, dv, dl, g_prof_buffer, dtype,
                               (cudaStream_t)stream, g_sparc_bwd_pdl_late);
    if (!sparc_tc_supported(P, T, D, dtype)) return CFA_ERR_WORKSPACE;
    if (g_split && q_save && sparc_bwd2_supported(P, T, D, dtype))      // streaming backward on the saved G / Q
      return sparc_bwd2_launch(v, l, mask, B, P, T, D, thr, scale, row_inv_norm, lse_row, lse_col, coef, tt_logits,
                               g_inv_norm, g_split, q_save, dpooled_v, dpooled_l, dv, dl, g_prof_buffer, dtype,
                               (cudaStream_t)stream);
    return sparc_bwd_tc_launch(v, l, mask, B, P, T, D, thr, scale, row_inv_norm, lse_row, lse_col, coef, tt_logits,
                               g_inv_norm, dpooled_v, dpooled_l, dv, dl, (cudaStream_t)stream);
  }
  return sparc_bwd_simt(v, l, mask, B, P, T, D, dtype, thr, scale, lse_row, lse_col, coef, dpooled_v, dpooled_l, dv, dl,
                        scratch, scratch_bytes, stream);
}
