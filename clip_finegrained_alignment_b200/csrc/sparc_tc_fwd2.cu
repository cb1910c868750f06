// Tensor-core SPARC forward, second generation (tcgen05 / TMEM / TMA) for bf16 embeddings on sm_100a.
//
// One CTA per sample: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2..9 = epilogue (two warps per TMEM lane
// quarter, each owning half of the columns; thread = token row).  Compared with sparc_fwd_tc_kernel:
//   - no separate prep kernel: while pass 0 streams the raw l / v tiles through shared memory, the epilogue warps
//     read the same tiles and accumulate the row norms (losses.py:221-222) and the pooled means (losses.py:207-212);
//   - twice the epilogue warps, shared-space LDS / STS, packed bf16 conversions, 3-deep TMA ring;
//   - saves G (bf16 hi | lo) and Q = G . v^T for the streaming backward (sparc_tc_bwd2.cu).
//
//   pass 0   S[T,P] = sum_kb l_kb . v_kb^T                   (TMEM cS)      + row norms / pooled means from smem
//   E1       min-max, threshold, renormalise -> W (hi/lo) in smem           (losses.py:228-243)
//   pass 1   G_kb = W . v_kb (v tile read MN-major)                         (losses.py:245)
//            epilogue: ||G||^2, G hi/lo -> smem operand + global
//            L[T,T] += G_kb . l_kb^T ;  Q[T,P] += G_kb . v_kb^T             (losses.py:180)
//   E3       masked row / column log-sum-exp and CE                         (losses.py:186-196)
#include "tc_common.cuh"
#include "sparc_paths.h"
#include <math_constants.h>

namespace cfa {
using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int kF2Threads = 352;      // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue, warp 10 pooled means
constexpr float kF2NormEps = 1e-12f, kF2MinMaxEps = 1e-8f, kF2ClampEps = 1e-8f;
constexpr uint32_t kF2cS = 0, kF2cG = 256, kF2cL = 384;

struct Fwd2Layout {
  int NP, NT, KB, NS;
  uint32_t l_bytes, v_bytes, slot, w_bytes, g_bytes;
  uint32_t off_w, off_g, off_f, off_x, off_bar, total;
};

__host__ __device__ inline Fwd2Layout fwd2_layout(int P, int T, int D, int NS) {
  Fwd2Layout L;
  L.NP = (P + 15) & ~15; L.NT = (T + 15) & ~15; L.KB = D / 64; L.NS = NS;
  L.l_bytes = L.NT * 128; L.v_bytes = L.NP * 128; L.slot = L.l_bytes + L.v_bytes;
  L.w_bytes = (uint32_t)L.NP * L.NT * 2;            // one of hi / lo, interleaved [NP/8][NT][8]
  L.g_bytes = 64u * L.NT * 2;                       // one of hi / lo of one G_kb operand buffer
  L.off_w = L.NS * L.slot;
  const uint32_t lb = (uint32_t)L.NT * (L.NT + 1) * 4;            // fp32 logits scratch aliases the W region
  L.off_g = L.off_w + (((2 * L.w_bytes > lb ? 2 * L.w_bytes : lb) + 1023) & ~1023u);
  L.off_f = L.off_g + 4 * L.g_bytes + 2048;         // phantom rows of the last interleaved chunk stay in bounds
  L.off_x = L.off_f + 4 * ((L.NP + 32) + 3 * L.NT + 32);
  L.off_bar = (L.off_x + 4 * (2 * 2 * L.NT * 4) + 7) & ~7u;
  L.total = L.off_bar + 8 * (3 * L.NS + 12) + 16;
  return L;
}

struct Fwd2Params {
  long long* prof;
  int P, T, D, NS;
  float thr, scale;
  const uint8_t* mask;
  float* inv_vn;        // [B][P]  out
  float* inv_ln;        // [B][T]  out
  float* pooled_v;      // [B][D]  out
  float* pooled_l;      // [B][D]  out
  float* lse_row;
  float* lse_col;
  float* local_partial;
  float* tt_logits;     // [B][T][T] masked, scaled logits (saved for the backward), may be NULL
  float* g_inv_norm;    // [B][T]
  bf16* g_split;        // [B][2][T][D], may be NULL
  float* q_save;        // [B][T][NP], may be NULL
};

// log-depth reductions of 16 register values (the epilogues run one or two warps per scheduler: a 16-long dependent
// chain of FADD / FMNMX costs more than the arithmetic it carries)
__device__ __forceinline__ float f2_max16(const float* x) {
  const float a = fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])), b = fmaxf(fmaxf(x[4], x[5]), fmaxf(x[6], x[7]));
  const float c = fmaxf(fmaxf(x[8], x[9]), fmaxf(x[10], x[11])), d = fmaxf(fmaxf(x[12], x[13]), fmaxf(x[14], x[15]));
  return fmaxf(fmaxf(a, b), fmaxf(c, d));
}
__device__ __forceinline__ float f2_sum16(const float* x) {
  const float a = (x[0] + x[1]) + (x[2] + x[3]), b = (x[4] + x[5]) + (x[6] + x[7]);
  const float c = (x[8] + x[9]) + (x[10] + x[11]), d = (x[12] + x[13]) + (x[14] + x[15]);
  return (a + b) + (c + d);
}

__device__ __forceinline__ void f2_epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Column sums of 16 per-lane values over the 32 lanes of a warp (31 shuffles): every lane gets the sum of v[lane & 15].
__device__ __forceinline__ float f2_colsum16(float* v, int lane) {
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], 16);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const bool up = lane & 8;
    const float send = up ? v[k] : v[k + 8], keep = up ? v[k + 8] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool up = lane & 4;
    const float send = up ? v[k] : v[k + 4], keep = up ? v[k + 4] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const bool up = lane & 2;
    const float send = up ? v[k] : v[k + 2], keep = up ? v[k + 2] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    const bool up = lane & 1;
    const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  return v[0];
}

// kHalf: the raw embeddings are fp16 (torch.autocast's default) instead of bf16; on-chip operands stay bf16 hi/lo
template <int kNksP, bool kHalf>
__global__ void __launch_bounds__(kF2Threads, 1)
sparc_fwd2_kernel(const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmL, const Fwd2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = CFA_SMEM_BASE_1024(smem_raw);
  const Fwd2Layout L = fwd2_layout(p.P, p.T, p.D, p.NS);
  const int NP = L.NP, NT = L.NT, KB = L.KB, NS = L.NS, P = p.P, T = p.T, D = p.D;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* Whi = base + L.off_w;
  uint8_t* Wlo = Whi + L.w_bytes;
  uint8_t* Gs = base + L.off_g;                       // [buf][hi|lo][g_bytes]
  float* ivn = (float*)(base + L.off_f);              // [NP + 32]: sum of squares, then 1 / max(|v_p|, eps); zero beyond P
  float* iln = ivn + NP + 32;                         // [NT]
  float* msk = iln + NT;                              // [NT]
  float* gnx = msk + NT;                              // [NT] spare
  float* red = gnx + NT;                              // [32]
  float* xch = (float*)(base + L.off_x);              // [2][2][NT][4]
  uint64_t* bars = (uint64_t*)(base + L.off_bar);
  uint64_t* full = bars;                              // [NS]
  uint64_t* empty0 = bars + NS;                       // [NS] pass-0 uses: MMA commit + 8 epilogue warps + pool warp
  uint64_t* empty1 = bars + 2 * NS;                   // [NS] pass-1 uses: MMA commit
  uint64_t* bb = bars + 3 * NS;
  uint64_t* s_full = bb + 0;
  uint64_t* w_ready = bb + 1;
  uint64_t* g_full = bb + 2;                          // [2]
  uint64_t* g_free = bb + 4;                          // [2]
  uint64_t* gs_ready = bb + 6;                        // [2]
  uint64_t* gs_free = bb + 8;                         // [2]
  uint64_t* l_full = bb + 10;
  uint32_t* tmem_slot = (uint32_t*)(bb + 11);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full + i, 1); mbar_init(empty0 + i, 10); mbar_init(empty1 + i, 1); }
    mbar_init(s_full, 1); mbar_init(w_ready, 8); mbar_init(l_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(g_full + i, 1); mbar_init(g_free + i, 8); mbar_init(gs_ready + i, 8); mbar_init(gs_free + i, 1); }
    fence_barrier_init();
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmL);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < NP + 32 + 2 * NT; i += kF2Threads) {
    if (i < NP + 32 + NT) ivn[i] = 0.f;               // ivn and iln start as sum-of-squares accumulators
    else { const int t = i - NP - 32 - NT; msk[t] = (t < T && p.mask[(size_t)b * T + t]) ? 1.f : 0.f; }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // slot use u: 0 .. KB-1 = pass 0, KB .. 2KB-1 = pass 1; slot = u % NS
  auto first1 = [&](int s) { const int r = KB % NS; return KB + ((s - r) % NS + NS) % NS; };   // first pass-1 use of slot s

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      for (int u = 0; u < 2 * KB; ++u) {
        const int s = u % NS, kb = u % KB;
        if (u >= NS) {                                // wait until the previous use of this slot has been consumed
          const int pu = u - NS;
          if (pu < KB) mbar_wait(empty0 + s, (pu / NS) & 1);
          else mbar_wait(empty1 + s, ((pu - first1(s)) / NS) & 1);
        }
        uint8_t* st = base + (size_t)s * L.slot;
        mbar_expect_tx(full + s, L.slot);
        tma_load_3d(st, &tmL, full + s, kb * 64, 0, b);
        tma_load_3d(st + L.l_bytes, &tmV, full + s, kb * 64, 0, b);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (warp-uniform control flow, one elected lane issues) ===============================
    const bool leader = elect_one();
    const uint32_t idesc_s = make_idesc16(128, NP, false, false, kHalf, kHalf);     // raw l x raw v
    const uint32_t idesc_g = make_idesc16(128, 64, false, true, false, kHalf);      // W (bf16 hi/lo) x raw v
    const uint32_t idesc_l = make_idesc16(128, NT, false, false, false, kHalf);     // G (bf16 hi/lo) x raw l
    const uint32_t idesc_q = make_idesc16(128, NP, false, false, false, kHalf);     // G (bf16 hi/lo) x raw v
    const uint32_t il_lbo = (uint32_t)NT * 16;
    const uint64_t sw0 = make_smem_desc(0, 16, 1024, kLayoutSw128);
    const uint64_t ilk_whi = make_smem_desc(smem_u32(Whi), il_lbo, 128, kLayoutNone);
    const uint64_t ilk_wlo = make_smem_desc(smem_u32(Wlo), il_lbo, 128, kLayoutNone);
    const uint32_t ilk_step = (2 * il_lbo) >> 4;
    long long* pf = (p.prof && leader) ? p.prof + (size_t)b * 32 : nullptr;
    int pi = 0;
    auto stamp = [&]() { if (pf) pf[pi++] = clock64(); };
    stamp();
    // ---- pass 0: S = l . v^T
    for (int u = 0; u < KB; ++u) {
      const int s = u % NS;
      mbar_wait(full + s, (u / NS) & 1);
      tc_fence_after();
      const uint32_t sl = smem_u32(base + (size_t)s * L.slot), sv = sl + L.l_bytes;
      const uint64_t dl0 = sw0 | (sl >> 4), dv0 = sw0 | (sv >> 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss_w(leader, tmem + kF2cS, dl0 + 2 * k, dv0 + 2 * k, idesc_s, (u | k) != 0);
      umma_commit_w(leader, empty0 + s);
    }
    umma_commit_w(leader, s_full);
    stamp();
    // ---- pass 1: G_kb = W . v_kb ; L += G_kb . l_kb^T ; Q += G_kb . v_kb^T   (G issued one block ahead)
    mbar_wait(w_ready, 0);
    tc_fence_after();
    stamp();
    const int nks = kNksP ? kNksP : NP / 16;
    long long wfull = 0, wgfree = 0, wgs = 0;
    auto issue_g = [&](int kb) {
      const int u = KB + kb, s = u % NS, buf = kb & 1;
      const long long w0 = clock64();
      mbar_wait(full + s, (u / NS) & 1);
      const long long w1 = clock64();
      mbar_wait(g_free + buf, ((kb >> 1) & 1) ^ 1);
      wfull += w1 - w0; wgfree += clock64() - w1;
      tc_fence_after();
      const uint64_t dv0 = sw0 | ((smem_u32(base + (size_t)s * L.slot) + L.l_bytes) >> 4);
      const uint32_t d = tmem + kF2cG + 64 * buf;
      _Pragma("unroll") for (int ks = 0; ks < nks; ++ks) umma_ss_w(leader, d, ilk_whi + ks * ilk_step, dv0 + ks * 128, idesc_g, ks != 0);
      _Pragma("unroll") for (int ks = 0; ks < nks; ++ks) umma_ss_w(leader, d, ilk_wlo + ks * ilk_step, dv0 + ks * 128, idesc_g, true);
      umma_commit_w(leader, g_full + buf);
    };
    auto issue_l = [&](int kb) {
      const int u = KB + kb, s = u % NS, buf = kb & 1;
      const long long w0 = clock64();
      mbar_wait(gs_ready + buf, (kb >> 1) & 1);
      wgs += clock64() - w0;
      tc_fence_after();
      const uint32_t sl = smem_u32(base + (size_t)s * L.slot);
      const uint64_t dl0 = sw0 | (sl >> 4), dv0 = sw0 | ((sl + L.l_bytes) >> 4);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint64_t ga = make_smem_desc(smem_u32(Gs + (size_t)(2 * buf + half) * L.g_bytes), il_lbo, 128, kLayoutNone);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss_w(leader, tmem + kF2cL, ga + k * ilk_step, dl0 + 2 * k, idesc_l, (kb | half | k) != 0);
        if (p.q_save) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss_w(leader, tmem + kF2cS, ga + k * ilk_step, dv0 + 2 * k, idesc_q, (kb | half | k) != 0);
        }
      }
      umma_commit_w(leader, gs_free + buf);
      umma_commit_w(leader, empty1 + s);
    };
    issue_g(0);
    for (int kb = 0; kb < KB; ++kb) {
      if (kb + 1 < KB) issue_g(kb + 1);
      issue_l(kb);
    }
    umma_commit_w(leader, l_full);
    stamp();
    if (pf) { pf[8] = wfull; pf[9] = wgfree; pf[10] = wgs; }
  } else if (warp == 10) {
    // =============================== pooled means (losses.py:207-212), pass 0 only ===============================
    // lane -> (16-byte chunk c = lane % 8, row group g = lane / 8): rows r = g, g+4, ... of the tile, 8 columns each
    // (conflict-free under the 128-byte swizzle, 20 independent 128-bit loads per block), then two shuffle steps over the
    // 4 row groups.  The text mean always comes from here; the image mean only when there is no spare MMA row (see E1).
    const bool pool_tc = T < NT;
    float cnt = 0.f;
    for (int t = 0; t < T; ++t) cnt += msk[t];
    const float inv_cnt = 1.f / fmaxf(cnt, kF2ClampEps), invP = 1.f / (float)P;
    const int c = lane & 7, g = lane >> 3;
    for (int u = 0; u < KB; ++u) {
      const int s = u % NS;
      mbar_wait(full + s, (u / NS) & 1);
      const uint8_t* st = base + (size_t)s * L.slot;
      float al[8], av[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) { al[i] = 0.f; av[i] = 0.f; }
#pragma unroll 5
      for (int r = g; r < NT; r += 4) {                 // rows beyond T are zero-filled by TMA and have msk = 0
        float f8[8];
        unpack_raw8<kHalf>(*reinterpret_cast<const uint4*>(st + r * 128 + ((c ^ (r & 7)) << 4)), f8);
        const float m = msk[r];
#pragma unroll
        for (int i = 0; i < 8; ++i) al[i] = fmaf(m, f8[i], al[i]);
      }
      if (!pool_tc) {
#pragma unroll 4
        for (int r = g; r < NP; r += 4) {
          float f8[8];
          unpack_raw8<kHalf>(*reinterpret_cast<const uint4*>(st + L.l_bytes + r * 128 + ((c ^ (r & 7)) << 4)), f8);
#pragma unroll
          for (int i = 0; i < 8; ++i) av[i] += f8[i];
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + s);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        al[i] += __shfl_xor_sync(0xffffffffu, al[i], 8);
        al[i] += __shfl_xor_sync(0xffffffffu, al[i], 16);
        if (!pool_tc) { av[i] += __shfl_xor_sync(0xffffffffu, av[i], 8); av[i] += __shfl_xor_sync(0xffffffffu, av[i], 16); }
      }
      if (g == 0) {
        float* pl = p.pooled_l + (size_t)b * D + u * 64 + 8 * c;
        *reinterpret_cast<float4*>(pl) = make_float4(al[0] * inv_cnt, al[1] * inv_cnt, al[2] * inv_cnt, al[3] * inv_cnt);
        *reinterpret_cast<float4*>(pl + 4) = make_float4(al[4] * inv_cnt, al[5] * inv_cnt, al[6] * inv_cnt, al[7] * inv_cnt);
        if (!pool_tc) {
          float* pv = p.pooled_v + (size_t)b * D + u * 64 + 8 * c;
          *reinterpret_cast<float4*>(pv) = make_float4(av[0] * invP, av[1] * invP, av[2] * invP, av[3] * invP);
          *reinterpret_cast<float4*>(pv + 4) = make_float4(av[4] * invP, av[5] * invP, av[6] * invP, av[7] * invP);
        }
      }
    }
  } else {
    // =============================== epilogue: 8 warps, 2 per TMEM lane quarter ===============================
    const int ew = warp - 2, q = warp & 3, h = ew >> 2;
    const int row = 32 * q + lane;
    const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
    const bool inT = row < NT;
    const int tid = threadIdx.x - 64;
    float* xme0 = xch + ((0 * 2 + h) * NT + (inT ? row : 0)) * 4;
    float* xot0 = xch + ((0 * 2 + (1 - h)) * NT + (inT ? row : 0)) * 4;
    const int xset = 2 * NT * 4;
    float cnt = 0.f;
    for (int t = 0; t < T; ++t) cnt += msk[t];
    const float inv_cnt = 1.f / fmaxf(cnt, kF2ClampEps), invP = 1.f / (float)P;
    long long* pf = (p.prof && row == 0 && h == 0) ? p.prof + (size_t)b * 32 + 16 : nullptr;
    int pi = 0;
    auto stamp = [&]() { if (pf) pf[pi++] = clock64(); };
    stamp();

    // ---- pass 0 side job: row norms and pooled means straight from the TMA tiles.  Warp ew owns the ew-th 16-byte chunk
    // (8 columns) of every row of the block, lane l the rows l, l+32, ...  (conflict-free under the 128-byte swizzle)
    const bool pool_tc = T < NT;                        // a spare MMA row exists (T not a multiple of 16): see E1
    float ssv[8], ssl[4];                               // this thread's rows: r = lane + 32 k  (NP <= 256, NT <= 128)
#pragma unroll
    for (int k = 0; k < 8; ++k) ssv[k] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) ssl[k] = 0.f;
    for (int u = 0; u < KB; ++u) {
      const int s = u % NS;
      mbar_wait(full + s, (u / NS) & 1);
      const uint8_t* st = base + (size_t)s * L.slot;
      // (the pooled IMAGE mean comes out of the tensor core — spare MMA row, see E1 — so this job stays short: a longer
      // hold on the pass-0 slots throttles the TMA stream)
#pragma unroll
      for (int k = 0; k < 8; ++k) {                     // v rows
        const int r = lane + 32 * k;
        if (r < NP) {
          const uint4 raw = *reinterpret_cast<const uint4*>(st + L.l_bytes + r * 128 + ((ew ^ (r & 7)) << 4));
          float f8[8];
          unpack_raw8<kHalf>(raw, f8);
          ssv[k] += (f8[0] * f8[0] + f8[1] * f8[1]) + (f8[2] * f8[2] + f8[3] * f8[3]) +
                    ((f8[4] * f8[4] + f8[5] * f8[5]) + (f8[6] * f8[6] + f8[7] * f8[7]));
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {                     // l rows (zero beyond T)
        const int r = lane + 32 * k;
        if (r < NT) {
          const uint4 raw = *reinterpret_cast<const uint4*>(st + r * 128 + ((ew ^ (r & 7)) << 4));
          float f8[8];
          unpack_raw8<kHalf>(raw, f8);
          ssl[k] += (f8[0] * f8[0] + f8[1] * f8[1]) + (f8[2] * f8[2] + f8[3] * f8[3]) +
                    ((f8[4] * f8[4] + f8[5] * f8[5]) + (f8[6] * f8[6] + f8[7] * f8[7]));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + s);           // this warp is done with the tile
    }
    stamp();
    {
      // each warp saw 8 of the 64 columns of every block: one cross-warp reduction (scratch = the still unused W region)
      float* part = reinterpret_cast<float*>(Whi);      // [8][NP + NT]
#pragma unroll
      for (int k = 0; k < 8; ++k) { const int r = lane + 32 * k; if (r < NP) part[ew * (NP + NT) + r] = ssv[k]; }
#pragma unroll
      for (int k = 0; k < 4; ++k) { const int r = lane + 32 * k; if (r < NT) part[ew * (NP + NT) + NP + r] = ssl[k]; }
      f2_epi_bar();
      for (int i = tid; i < NP + NT; i += 256) {
        float ss = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) ss += part[w * (NP + NT) + i];
        if (i < NP) {
          const float n = (i < P) ? 1.f / fmaxf(sqrtf(ss), kF2NormEps) : 0.f;
          ivn[i] = n;
          if (i < P) p.inv_vn[(size_t)b * P + i] = n;
        } else {
          const int t = i - NP;
          const float n = (t < T) ? 1.f / fmaxf(sqrtf(ss), kF2NormEps) : 0.f;
          iln[t] = n;
          if (t < T) p.inv_ln[(size_t)b * T + t] = n;
        }
      }
      f2_epi_bar();
    }
    const bool valid = row < T && msk[inT ? row : 0] != 0.f;
    const float il = inT ? iln[row] : 0.f;
    const bool pool_row = pool_tc && row == T;

    // ---- E1: S -> W   (column halves; row statistics exchanged through shared memory)
    const int psplit = ((NP / 16 + 1) / 2) * 16;
    const int c_lo = h ? psplit : 0, c_hi = h ? NP : psplit;
    mbar_wait(s_full, 0);
    tc_fence_after();
    stamp();
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
      float x[16], lo16[16], hi16[16];
      tmem_ld16(trow + kF2cS + c0, x);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float s = x[j] * il * ivn[c0 + j];
        const bool in = c0 + j < P;
        lo16[j] = in ? -s : -CUDART_INF_F;             // min via max of the negated values
        hi16[j] = in ? s : -CUDART_INF_F;
      }
      mn = fminf(mn, -f2_max16(lo16));
      mx = fmaxf(mx, f2_max16(hi16));
    }
    {
      float* xme = xme0 + xset;
      float* xot = xot0 + xset;
      if (inT) { xme[0] = mn; xme[1] = mx; }
      f2_epi_bar();
      if (inT) { mn = fminf(mn, xot[0]); mx = fmaxf(mx, xot[1]); }
    }
    const float inv_rng = 1.f / (mx - mn + kF2MinMaxEps);
    float sum = 0.f;
    for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
      float x[16];
      tmem_ld16(trow + kF2cS + c0, x);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float nn = (x[j] * il * ivn[c0 + j] - mn) * inv_rng;
        x[j] = (c0 + j < P && !(nn < p.thr)) ? nn : 0.f;
      }
      sum += f2_sum16(x);
    }
    if (inT) xme0[0] = sum;
    f2_epi_bar();
    if (inT) sum += xot0[0];
    const float inv_sigma = valid ? 1.f / fmaxf(sum, kF2ClampEps) : 0.f;
    for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
      float x[16];
      tmem_ld16(trow + kF2cS + c0, x);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float nn = (x[j] * il * ivn[c0 + j] - mn) * inv_rng;
        x[j] = (valid && c0 + j < P && !(nn < p.thr)) ? nn * inv_sigma : 0.f;
        // the first unused MMA row (t = T < NT) carries 1/P: G[T] = W[T] . v is then the pooled image embedding
        // (losses.py:207) for free — the tensor core computes all 128 rows anyway
        if (pool_row) x[j] = (c0 + j < P) ? invP : 0.f;
      }
      if (inT) {
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          uint4 hi, lo;
          split_hilo8(x + 8 * g8, hi, lo);
          const uint32_t off = il_offset(NT, row, c0 + 8 * g8);
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

    // ---- pass 1 epilogue: G_kb -> ||G||^2, bf16 hi/lo A operand (smem) + saved copy (global); 32 of the 64 columns per warp
    float gn2 = 0.f;
    for (int kb = 0; kb < KB; ++kb) {
      const int buf = kb & 1;
      mbar_wait(g_full + buf, (kb >> 1) & 1);
      tc_fence_after();
      float x[32];
      tmem_ld32(trow + kF2cG + 64 * buf + 32 * h, x);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(g_free + buf);
      mbar_wait(gs_free + buf, ((kb >> 1) & 1) ^ 1);
      if (pool_row) {                                   // G[T] = mean_p v[p] for this warp's 32 columns
        float* pv = p.pooled_v + (size_t)b * D + kb * 64 + 32 * h;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(pv + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
      }
      if (inT) {
        uint8_t* gh = Gs + (size_t)(2 * buf) * L.g_bytes;
        uint8_t* gl = gh + L.g_bytes;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float y[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { y[j] = valid ? x[8 * g + j] : 0.f; gn2 = fmaf(y[j], y[j], gn2); }
          uint4 hi, lo;
          split_hilo8(y, hi, lo);
          const uint32_t off = il_offset(NT, row, 32 * h + 8 * g);
          *reinterpret_cast<uint4*>(gh + off) = hi;
          *reinterpret_cast<uint4*>(gl + off) = lo;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(gs_ready + buf);
      // saved copy for the backward, OFF the critical chain and transposed: this warp re-reads the 32 rows x 4 chunks it
      // has just written (lane -> row 8 it + lane / 4, chunk lane % 4), so one STG covers 8 rows x 64 contiguous bytes
      // instead of 32 rows x 16 bytes.  The buffer is only overwritten by this same warp two blocks later.
      if (p.g_split) {
        const uint8_t* gh = Gs + (size_t)(2 * buf) * L.g_bytes;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int r = 32 * q + it * 8 + (lane >> 2), c = 4 * h + (lane & 3);
          if (r < T) {
            const uint32_t off = il_offset(NT, r, 8 * c);
            const uint4 hi = *reinterpret_cast<const uint4*>(gh + off), lo = *reinterpret_cast<const uint4*>(gh + L.g_bytes + off);
            bf16* gdst = p.g_split + (((size_t)b * 2) * T + r) * D + kb * 64 + 8 * c;
            *reinterpret_cast<uint4*>(gdst) = hi;
            *reinterpret_cast<uint4*>(gdst + (size_t)T * D) = lo;
          }
        }
      }
    }
    {
      float* xme = xme0 + xset;
      float* xot = xot0 + xset;
      if (inT) xme[0] = gn2;
      f2_epi_bar();
      if (inT) gn2 += xot[0];
    }
    const float ign = 1.f / fmaxf(sqrtf(gn2), kF2NormEps);
    stamp();

    // ---- E3: masked logits.  Half 0: row pass (LSE in registers, logits -> smem); half 1: Q -> global meanwhile.
    mbar_wait(l_full, 0);
    tc_fence_after();
    stamp();
    float* Lb = reinterpret_cast<float*>(Whi);          // [T][NT+1]  (the W region is free now)
    const int ldl = NT + 1;
    // row pass: each half owns part of the columns; (max, sum exp) pairs are merged through shared memory
    const int tsplit = ((NT / 16 + 1) / 2) * 16;
    const int t_lo = h ? tsplit : 0, t_hi = h ? NT : tsplit;
    const float sc_row = valid ? p.scale * ign : 0.f;
    float rmax = -CUDART_INF_F, rsum = 0.f, diag = 0.f;
    for (int c0 = t_lo; c0 < t_hi; c0 += 16) {
      float x[16];
      tmem_ld16(trow + kF2cL + c0, x);
      tmem_ld_wait();
#pragma unroll
      for (int j4 = 0; j4 < 16; j4 += 4) {
        const float4 mk = *reinterpret_cast<const float4*>(msk + c0 + j4), ik = *reinterpret_cast<const float4*>(iln + c0 + j4);
        const float mv[4] = {mk.x, mk.y, mk.z, mk.w}, iv[4] = {ik.x, ik.y, ik.z, ik.w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = j4 + jj, col = c0 + j;
          x[j] = (valid && mv[jj] != 0.f) ? x[j] * sc_row * iv[jj] : -CUDART_INF_F;   // msk/iln are 0 beyond T
          if (row < T) Lb[row * ldl + col] = x[j];
          diag = (col == row) ? x[j] : diag;
        }
      }
      const float nm = fmaxf(rmax, f2_max16(x));
      float ex[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) ex[j] = (x[j] == -CUDART_INF_F) ? 0.f : __expf(x[j] - nm);
      rsum = rsum * ((rmax == -CUDART_INF_F) ? 0.f : __expf(rmax - nm)) + f2_sum16(ex);
      rmax = nm;
    }
    if (inT) { xme0[0] = rmax; xme0[1] = rsum; xme0[2] = diag; }
    stamp();
    f2_epi_bar();                                       // exchange set 0; the logits are in smem
    stamp();
    float ce = 0.f;
    if (h == 0) {
      if (valid) {
        const float om = xot0[0], os = xot0[1];
        const float M = fmaxf(rmax, om);
        const float ssum = rsum * __expf(rmax - M) + ((om == -CUDART_INF_F) ? 0.f : os * __expf(om - M));
        const float lse = M + logf(ssum);
        p.lse_row[(size_t)b * T + row] = lse;
        ce = lse - ((row < tsplit) ? diag : xot0[2]);
      } else if (row < T) {
        p.lse_row[(size_t)b * T + row] = 0.f;
      }
      if (p.g_inv_norm && row < T) p.g_inv_norm[(size_t)b * T + row] = ign;
    }
    // column pass: thread `row` of half h covers rows [i_lo, i_hi) of column `row`
    {
      const int isplit = (T + 1) / 2;
      const int i_lo = h ? isplit : 0, i_hi = h ? T : isplit;
      float cmax = -CUDART_INF_F, csum = 0.f;
      if (valid) {                                       // 4 independent chains per loop
        float m4[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F}, s4[4] = {0.f, 0.f, 0.f, 0.f};
        int i = i_lo;
        for (; i + 3 < i_hi; i += 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) m4[k] = fmaxf(m4[k], Lb[(i + k) * ldl + row]);
        }
        for (; i < i_hi; ++i) m4[0] = fmaxf(m4[0], Lb[i * ldl + row]);
        cmax = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        i = i_lo;
        for (; i + 3 < i_hi; i += 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) { const float y = Lb[(i + k) * ldl + row]; s4[k] += (y == -CUDART_INF_F) ? 0.f : __expf(y - cmax); }
        }
        for (; i < i_hi; ++i) { const float y = Lb[i * ldl + row]; s4[0] += (y == -CUDART_INF_F) ? 0.f : __expf(y - cmax); }
        csum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      }
      float* xme = xme0 + xset;
      float* xot = xot0 + xset;
      if (inT) { xme[0] = cmax; xme[1] = csum; }
      stamp();
      f2_epi_bar();                                     // exchange set 1
      stamp();
      if (h == 1) {
        if (valid) {
          const float om = xot[0], os = xot[1];
          const float M = fmaxf(cmax, om);
          const float ssum = ((cmax == -CUDART_INF_F) ? 0.f : csum * __expf(cmax - M)) + ((om == -CUDART_INF_F) ? 0.f : os * __expf(om - M));
          const float lse = M + logf(ssum);
          p.lse_col[(size_t)b * T + row] = lse;
          ce = lse - Lb[row * ldl + row];
        } else if (row < T) {
          p.lse_col[(size_t)b * T + row] = 0.f;
        }
      }
    }
    if (p.q_save) {       // Q (TMEM columns of the dead S) -> global, row stride NP; column halves; transposed through a
      // per-warp smem tile (behind the logits scratch in the W region) so that one STG covers 8 rows x 64 bytes
      float* tile = reinterpret_cast<float*>(Whi + (((uint32_t)NT * (NT + 1) * 4 + 127) & ~127u)) + ew * (32 * 20);
      for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
        float x[16];
        tmem_ld16(trow + kF2cS + c0, x);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(tile + lane * 20 + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
        __syncwarp();
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rl = it * 8 + (lane >> 2), r = 32 * q + rl;
          const float4 v4 = *reinterpret_cast<const float4*>(tile + rl * 20 + 4 * (lane & 3));
          if (r < T) *reinterpret_cast<float4*>(p.q_save + ((size_t)b * T + r) * NP + c0 + 4 * (lane & 3)) = v4;
        }
        __syncwarp();
      }
    }
    stamp();
    if (p.tt_logits) {                                  // coalesced copy of the T x T logits for the backward
      float* dst = p.tt_logits + (size_t)b * T * T;
      for (int i = ew; i < T; i += 8)
        for (int j = lane; j < T; j += 32) dst[i * T + j] = Lb[i * ldl + j];
    }
    ce = warp_sum(ce);
    if (lane == 0) red[ew] = ce;                        // warps 0..3 (half 0): row direction, 4..7: column direction
    f2_epi_bar();
    if (tid == 0) {
      p.local_partial[2 * b] = red[0] + red[1] + red[2] + red[3];
      p.local_partial[2 * b + 1] = red[4] + red[5] + red[6] + red[7];
    }
    stamp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

static int fwd2_pick_stages(int P, int T, int D) {
  for (int ns = 4; ns >= 2; --ns)
    if (fwd2_layout(P, T, D, ns).total + 1024 <= 227 * 1024) return ns;
  return 0;
}

bool sparc_fwd2_supported(int P, int T, int D, int dtype) {
  // bf16 only: tcgen05 kind::f16 rejects MIXED operand formats (an fp16 raw tile against the bf16 hi/lo operands produced
  // on chip raises "illegal instruction" on sm_100a -- measured).  fp16 embeddings run on the third-generation kernels
  // (sparc_tc_fwd3.cu / sparc_tc_bwd3.cu), whose on-chip operands are range-scaled fp16 hi|lo; this generation's kHalf
  // template parameter is never instantiated.
  if (dtype != CFA_DTYPE_BF16) return false;
  if (!sparc_tc_supported(P, T, D, CFA_DTYPE_BF16)) return false;           // shape limits of the tensor-core layouts
  return fwd2_pick_stages(P, T, D) != 0;
}

int sparc_fwd2_launch(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, float thr,
                      float scale, float* row_inv_norm, float* pooled_v, float* pooled_l, float* lse_row, float* lse_col,
                      float* local_partial, float* tt_logits, float* g_inv_norm, void* g_split, float* q_save,
                      long long* prof, int dtype, cudaStream_t st) {
  const bool half = dtype == CFA_DTYPE_F16;
  const int NS = fwd2_pick_stages(P, T, D);
  if (NS == 0) return CFA_ERR_UNSUPPORTED;
  const Fwd2Layout L = fwd2_layout(P, T, D, NS);
  CUtensorMap tmV, tmL;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tmV, v, D, P, B, 64, L.NP, half)) != CFA_OK) return rc;
  if ((rc = make_tmap_bf16_3d(&tmL, l, D, T, B, 64, L.NT, half)) != CFA_OK) return rc;
  Fwd2Params prm{prof, P, T, D, NS, thr, scale, mask, row_inv_norm, row_inv_norm + (size_t)B * P, pooled_v, pooled_l, lse_row,
                 lse_col, local_partial, tt_logits, g_inv_norm, (bf16*)g_split, q_save};
  const size_t smem = L.total + 1024;
#define CFA_F2_LAUNCH(NKS, HALF)                                                                                          \
  do {                                                                                                                    \
    CFA_CUDA_TRY(cudaFuncSetAttribute(sparc_fwd2_kernel<NKS, HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    sparc_fwd2_kernel<NKS, HALF><<<B, kF2Threads, smem, st>>>(tmV, tmL, prm);                                             \
  } while (0)
  if (half) return CFA_ERR_UNSUPPORTED;
  if (L.NP == 208) CFA_F2_LAUNCH(13, false); else CFA_F2_LAUNCH(0, false);
#undef CFA_F2_LAUNCH
  return launch_status();
}

}  // namespace cfa
