// Tensor-core SPARC forward, third generation ("transposed" orientation) for bf16 embeddings on sm_100a.
//
// Everything the tensor core multiplies with a RAW embedding tile has the raw tile as the A operand (M = patches or
// feature columns: all 128 MMA rows live) and the operand produced on chip as the B operand, its bf16 hi and lo halves
// STACKED along N (N = 2 x 80 = 160): one tcgen05.mma per k-step delivers the hi and the lo product in adjacent TMEM
// column ranges, is tensor-bound (80 cycles measured, 79.3) instead of shared-memory-bound (N = 64: 47 cycles for 32),
// and the 77 -> 128 row padding of the token-major orientation disappears.
//
//   P0   S^T[p,t]  = sum_kb v_kb . l_kb^T          (A = v tile K-major, B = l tile; 64-wide D blocks)    losses.py:225
//        side jobs on the same tiles: row norms (losses.py:221-222), pooled text mean (losses.py:210-212)
//   E1   thread = patch p: cross-lane min / max per token (CREDUX), threshold -> Theta^T (un-normalised weights, hi|lo)
//        and S_raw^T (hi|lo) as [t/8][p][t%8] operands; sigma_t = sum_p Theta by shuffle transposition   losses.py:228-243
//   P1   L'^T[j,t] = S_raw . Theta^T               (K = p; replaces G . l^T: L_raw = W . S_raw^T)         losses.py:180
//        G'^T[d,t] = v^T . Theta^T per 128-wide D block (A = v tile pair read MN-major), epilogue: / sigma_t,
//        ||G_t||^2, G -> global as bf16 hi | lo planes (the backward's TMA operands); spare column t = T carries
//        1/P: the pooled image mean (losses.py:207)                                                    losses.py:245
//   E3   logits = s L' / (sigma_t ||G_t|| ||l_j||), masked row / column log-sum-exp, CE                 losses.py:186-196
#include "tc_common.cuh"
#include "sparc_paths.h"
#include <math_constants.h>

namespace cfa {
using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int kF3EpiWarps = 16;
constexpr int kF3Threads = 32 * (2 + kF3EpiWarps);   // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue
constexpr float kF3NormEps = 1e-12f, kF3MinMaxEps = 1e-8f, kF3ClampEps = 1e-8f;

struct Fwd3Layout {
  int NP, NT, MB, KB0, NBLK, CR0, NCH, NS0, NS1;
  uint32_t v_bytes, l_bytes, slot0, slot1, plane;      // plane = one of hi / lo of an interleaved [NT/8][NP][8] operand
  uint32_t off_th, off_sr, off_ring1, off_f, off_bar, total;
};

__host__ __device__ inline Fwd3Layout fwd3_layout(int P, int T, int D) {
  Fwd3Layout L;
  L.NP = (P + 15) & ~15; L.NT = (T + 15) & ~15; L.MB = L.NP > 128 ? 2 : 1; L.KB0 = D / 64; L.NBLK = D / 128;
  L.NCH = L.NP > 128 ? 2 : 1;
  L.CR0 = L.NCH == 2 ? 16 * ((L.NP + 31) / 32) : L.NP;
  L.v_bytes = (uint32_t)L.NP * 128; L.l_bytes = (uint32_t)L.NT * 128; L.slot0 = L.v_bytes + L.l_bytes;
  L.slot1 = 2u * L.CR0 * 128;
  L.plane = (uint32_t)L.NT * L.NP * 2;
  // the Theta^T region doubles as the P0 scratch: row-norm partials [8][NP + NT] and pooled-mean partials [4][D]
  const uint32_t scratch = (8u * (L.NP + L.NT) + 4u * D) * 4;
  const uint32_t op = ((2 * L.plane > scratch ? 2 * L.plane : scratch) + 1023) & ~1023u;
  L.off_th = 0; L.off_sr = op; L.off_ring1 = 2 * op;
  const uint32_t nf = (uint32_t)L.NP + 7 * L.NT + 8 * 2 * L.NT + 8 * L.NT + 64;
  const uint32_t fixed = 4 * nf + 8 * 32 + 1024;
  const uint32_t budget = 227u * 1024u;
  L.NS1 = 0; L.NS0 = 0; L.off_f = 0; L.off_bar = 0; L.total = 0;
  if (L.off_ring1 + fixed + 2 * L.slot1 > budget) return L;
  L.NS1 = (budget - L.off_ring1 - fixed) / L.slot1 >= 3 ? 3 : 2;
  uint32_t data_end = L.off_ring1 + L.NS1 * L.slot1;
  // the P0 ring aliases the S_raw^T and P1-ring regions (both dead during P0); M block 1 of the last slot reads 128 rows
  // past row 128 of its v tile, which must stay inside the allocation
  int ns0 = (int)((data_end - L.off_sr) / L.slot0);
  L.NS0 = ns0 > 4 ? 4 : ns0;
  if (L.NS0 >= 1) {
    const uint32_t reach = L.off_sr + (L.NS0 - 1) * L.slot0 + 256u * 128u;
    if (reach > data_end) data_end = reach;
  }
  L.off_f = (data_end + 127) & ~127u;
  L.off_bar = (L.off_f + 4 * nf + 7) & ~7u;
  L.total = L.off_bar + 8 * 32;
  return L;
}

struct Fwd3Params {
  long long* prof;
  int P, T, D;
  float thr, scale;
  const uint8_t* mask;
  float* inv_vn;        // [B][P]  out
  float* inv_ln;        // [B][T]  out
  float* pooled_v;      // [B][D]  out
  float* pooled_l;      // [B][D]  out
  float* lse_row;
  float* lse_col;
  float* local_partial;
  float* tt_logits;     // [B][T][T] masked, scaled logits (saved for the backward)
  float* g_inv_norm;    // [B][T]
  bf16* g_split;        // [B][2][T][D] hi | lo planes
  float* stats;         // [B][T][4]: min, 1 / (max - min + eps), sigma, arg-min patch (int bits)
};

__device__ __forceinline__ void f3_epi_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

template <bool kHalf>
__global__ void __launch_bounds__(kF3Threads, 1)
sparc_fwd3_kernel(const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmV1,
                  const __grid_constant__ CUtensorMap tmL, const Fwd3Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = CFA_SMEM_BASE_1024(smem_raw);
  const Fwd3Layout L = fwd3_layout(p.P, p.T, p.D);
  const int NP = L.NP, NT = L.NT, MB = L.MB, KB0 = L.KB0, NS0 = L.NS0, NS1 = L.NS1, P = p.P, T = p.T, D = p.D;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* TH = base + L.off_th;                      // Theta^T  [2 NT / 8][NP][8]: hi chunks, then lo chunks
  uint8_t* SR = base + L.off_sr;                      // S_raw^T  same layout
  uint8_t* ring0 = base + L.off_sr;                   // P0 slots [v tile | l tile]
  uint8_t* ring1 = base + L.off_ring1;                // P1 slots [v chunk, d 0..63 | v chunk, d 64..127]
  float* ivn = (float*)(base + L.off_f);              // [NP]
  float* iln = ivn + NP;                              // [NT]
  float* msk = iln + NT;                              // [NT]
  float* mnf = msk + NT;                              // [NT] row minimum
  float* irf = mnf + NT;                              // [NT] 1 / (max - min + eps)
  float* isg = irf + NT;                              // [NT] 1 / sigma (0 for masked tokens)
  float* ign = isg + NT;                              // [NT] 1 / ||G_t||
  int* imn = (int*)(ign + NT);                        // [NT] arg-min patch
  float* part_mm = (float*)(imn + NT);                // [8][NT][2]
  float* part_s = part_mm + 8 * 2 * NT;               // [8][NT]   (later: ||G||^2 partials [4][NT])
  float* red = part_s + 8 * NT;                       // [64]
  uint64_t* bars = (uint64_t*)(base + L.off_bar);
  uint64_t* full0 = bars;                             // [4]
  uint64_t* empty0 = bars + 4;                        // [4] MMA commit + 16 epilogue warps
  uint64_t* full1 = bars + 8;                         // [3]
  uint64_t* empty1 = bars + 11;                       // [3]
  uint64_t* s_full = bars + 14;
  uint64_t* w_ready = bars + 15;
  uint64_t* l_full = bars + 16;
  uint64_t* g_full = bars + 17;                       // [2]
  uint64_t* g_free = bars + 19;                       // [2]
  uint32_t* tmem_slot = (uint32_t*)(bars + 21);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) { mbar_init(full0 + i, 1); mbar_init(empty0 + i, 1 + kF3EpiWarps); }
    for (int i = 0; i < 3; ++i) { mbar_init(full1 + i, 1); mbar_init(empty1 + i, 1); }
    mbar_init(s_full, 1); mbar_init(w_ready, kF3EpiWarps); mbar_init(l_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(g_full + i, 1); mbar_init(g_free + i, kF3EpiWarps); }
    fence_barrier_init();
    tma_prefetch_desc(&tmV0); tma_prefetch_desc(&tmV1); tma_prefetch_desc(&tmL);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < NT; i += kF3Threads) {
    msk[i] = (i < T && p.mask[(size_t)b * T + i]) ? 1.f : 0.f;
    imn[i] = 0x7fffffff;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int NT2 = 2 * NT;
  const uint32_t cG = (uint32_t)NT2;                  // G'^T ping-pong buffers at columns NT2 and 2 NT2

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      for (int u = 0; u < KB0; ++u) {
        const int s = u % NS0;
        if (u >= NS0) mbar_wait(empty0 + s, ((u / NS0) - 1) & 1);
        uint8_t* st = ring0 + (size_t)s * L.slot0;
        mbar_expect_tx(full0 + s, L.slot0);
        tma_load_3d(st, &tmV0, full0 + s, u * 64, 0, b);
        tma_load_3d(st + L.v_bytes, &tmL, full0 + s, u * 64, 0, b);
      }
      // the P1 ring overlaps the P0 slots: every P0 use must have been released
      for (int s = 0; s < NS0 && s < KB0; ++s) {
        const int n_s = (KB0 - s + NS0 - 1) / NS0;
        mbar_wait(empty0 + s, (n_s - 1) & 1);
      }
      const int n1 = L.NBLK * L.NCH;
      for (int i = 0; i < n1; ++i) {
        const int s = i % NS1, blk = i / L.NCH, ch = i % L.NCH;
        if (i >= NS1) mbar_wait(empty1 + s, ((i / NS1) - 1) & 1);
        uint8_t* st = ring1 + (size_t)s * L.slot1;
        mbar_expect_tx(full1 + s, L.slot1);
        tma_load_3d(st, &tmV1, full1 + s, blk * 128, ch * L.CR0, b);
        tma_load_3d(st + L.slot1 / 2, &tmV1, full1 + s, blk * 128 + 64, ch * L.CR0, b);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (warp-uniform control flow, one elected lane issues) ===============================
    const bool leader = elect_one();
    const uint32_t idesc_s = make_idesc16(128, NT, false, false, kHalf, kHalf);     // raw v (K-major) x raw l (K-major)
    const uint32_t idesc_l2 = make_idesc16(128, NT2, true, true, false, false);     // S_raw^T (MN-major) x Theta^T hi|lo (MN-major)
    const uint32_t idesc_l1 = make_idesc16(128, NT, true, true, false, false);
    const uint32_t idesc_g = make_idesc16(128, NT2, true, true, kHalf, false);      // raw v^T (MN-major tile pair) x Theta^T hi|lo
    const uint64_t sw0 = make_smem_desc(0, 16, 1024, kLayoutSw128);
    long long* pf = (p.prof && leader) ? p.prof + (size_t)b * 32 : nullptr;
    int pi = 0;
    auto stamp = [&]() { if (pf) pf[pi++] = clock64(); };
    stamp();
    // ---- P0: S^T = v . l^T
    for (int u = 0; u < KB0; ++u) {
      const int s = u % NS0;
      mbar_wait(full0 + s, (u / NS0) & 1);
      tc_fence_after();
      const uint32_t sv = smem_u32(ring0 + (size_t)s * L.slot0), sl = sv + L.v_bytes;
      const uint64_t dv0 = sw0 | (sv >> 4), dl0 = sw0 | (sl >> 4);
      for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_ss_w(leader, tmem + mb * NT, dv0 + mb * (16384 >> 4) + 2 * k, dl0 + 2 * k, idesc_s, (u | k) != 0);
      }
      umma_commit_w(leader, empty0 + s);
    }
    umma_commit_w(leader, s_full);
    stamp();
    // ---- P1: L'^T = S_raw . Theta^T (into the dead S^T columns), then G'^T per 128-wide D block
    mbar_wait(w_ready, 0);
    tc_fence_after();
    stamp();
    const uint32_t il_sbo = (uint32_t)NP * 16;
    const uint64_t m_th = make_smem_desc(smem_u32(TH), 128, il_sbo, kLayoutNone);
    const uint64_t m_srh = make_smem_desc(smem_u32(SR), 128, il_sbo, kLayoutNone);
    const uint64_t m_srl = make_smem_desc(smem_u32(SR) + L.plane, 128, il_sbo, kLayoutNone);
    const int nksP = NP / 16;
    for (int ks = 0; ks < nksP; ++ks) umma_ss_w(leader, tmem, m_srh + ks * 16, m_th + ks * 16, idesc_l2, ks != 0);
    for (int ks = 0; ks < nksP; ++ks) umma_ss_w(leader, tmem, m_srl + ks * 16, m_th + ks * 16, idesc_l1, true);
    umma_commit_w(leader, l_full);
    int i1 = 0;
    for (int blk = 0; blk < L.NBLK; ++blk) {
      const int buf = blk & 1;
      mbar_wait(g_free + buf, ((blk >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d = tmem + cG + buf * NT2;
      for (int ch = 0; ch < L.NCH; ++ch, ++i1) {
        const int s = i1 % NS1;
        mbar_wait(full1 + s, (i1 / NS1) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(ring1 + (size_t)s * L.slot1);
        const uint64_t da = make_smem_desc(sa, L.slot1 / 2, 1024, kLayoutSw128);
        const int r0 = ch * L.CR0, nk = (ch == 0 ? L.CR0 : NP - L.CR0) / 16;
        const uint64_t db = m_th + (uint32_t)r0;                      // r0 rows x 16 B, in 16-byte units
        for (int ks = 0; ks < nk; ++ks) umma_ss_w(leader, d, da + ks * 128, db + ks * 16, idesc_g, (ch | ks) != 0);
        umma_commit_w(leader, empty1 + s);
      }
      umma_commit_w(leader, g_full + buf);
    }
    stamp();
  } else {
    // =============================== epilogue: 16 warps = 4 TMEM lane quarters x 4 groups ===============================
    const int ew = warp - 2, q = warp & 3, grp = ew >> 2;
    const uint32_t tq = tmem + ((uint32_t)(32 * q) << 16);
    const int tid = ew * 32 + lane;                     // 0..511
    const bool pool_tc = T < NT;
    long long* pf = (p.prof && tid == 0) ? p.prof + (size_t)b * 32 + 16 : nullptr;
    int pi = 0;
    auto stamp = [&]() { if (pf) pf[pi++] = clock64(); };
    stamp();

    // ---- P0 side job: row sums of squares (losses.py:221-222) and the pooled means (losses.py:207-212) straight from the
    // TMA tiles.  Warp -> 16-byte chunk c (8 columns of the 64-wide block) and row parity: rows lane + 32 k, k = hf, hf + 2,
    // ... (conflict-free under the 128-byte swizzle).  The pooled IMAGE mean comes from here only when there is no spare
    // token column (T == NT), see E1.
    {
      const int c = ew & 7, hf = ew >> 3;
      float* part = reinterpret_cast<float*>(TH);       // [8 chunks][NP + NT] (the Theta region is unused until E1)
      float* poolp = part + 8 * (NP + NT);              // [2 modalities][2 parities][D]
      float ssv[4] = {0.f, 0.f, 0.f, 0.f}, ssl[2] = {0.f, 0.f};
      for (int u = 0; u < KB0; ++u) {
        const int s = u % NS0;
        mbar_wait(full0 + s, (u / NS0) & 1);
        const uint8_t* st = ring0 + (size_t)s * L.slot0;
        float al[8], av[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { al[i] = 0.f; av[i] = 0.f; }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int r = lane + 32 * (2 * k + hf);
          if (r < NP) {
            float f8[8];
            unpack_raw8<kHalf>(*reinterpret_cast<const uint4*>(st + r * 128 + ((c ^ (r & 7)) << 4)), f8);
            ssv[k] += (f8[0] * f8[0] + f8[1] * f8[1]) + (f8[2] * f8[2] + f8[3] * f8[3]) +
                      ((f8[4] * f8[4] + f8[5] * f8[5]) + (f8[6] * f8[6] + f8[7] * f8[7]));
            if (!pool_tc) {
#pragma unroll
              for (int i = 0; i < 8; ++i) av[i] += f8[i];                // rows beyond P are zero-filled by TMA
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int r = lane + 32 * (2 * k + hf);
          if (r < NT) {
            float f8[8];
            unpack_raw8<kHalf>(*reinterpret_cast<const uint4*>(st + L.v_bytes + r * 128 + ((c ^ (r & 7)) << 4)), f8);
            ssl[k] += (f8[0] * f8[0] + f8[1] * f8[1]) + (f8[2] * f8[2] + f8[3] * f8[3]) +
                      ((f8[4] * f8[4] + f8[5] * f8[5]) + (f8[6] * f8[6] + f8[7] * f8[7]));
            const float m = msk[r];
#pragma unroll
            for (int i = 0; i < 8; ++i) al[i] = fmaf(m, f8[i], al[i]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + s);
        const float sl = warp_colsum8(al, lane);         // lane 2 i: column i of this chunk, summed over the warp's rows
        if ((lane & 17) == 0) poolp[(2 + hf) * D + u * 64 + 8 * c + (lane >> 1)] = sl;
        if (!pool_tc) {
          const float sv = warp_colsum8(av, lane);
          if ((lane & 17) == 0) poolp[hf * D + u * 64 + 8 * c + (lane >> 1)] = sv;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) { const int r = lane + 32 * (2 * k + hf); if (r < NP) part[c * (NP + NT) + r] = ssv[k]; }
#pragma unroll
      for (int k = 0; k < 2; ++k) { const int r = lane + 32 * (2 * k + hf); if (r < NT) part[c * (NP + NT) + NP + r] = ssl[k]; }
      f3_epi_bar();
      for (int i = tid; i < NP + NT; i += 512) {
        float ss = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) ss += part[w * (NP + NT) + i];
        if (i < NP) {
          const float n = (i < P) ? 1.f / fmaxf(sqrtf(ss), kF3NormEps) : 0.f;
          ivn[i] = n;
          if (i < P) p.inv_vn[(size_t)b * P + i] = n;
        } else {
          const int t = i - NP;
          const float n = (t < T) ? 1.f / fmaxf(sqrtf(ss), kF3NormEps) : 0.f;
          iln[t] = n;
          if (t < T) p.inv_ln[(size_t)b * T + t] = n;
        }
      }
      {
        float cnt = 0.f;
        for (int t = 0; t < T; ++t) cnt += msk[t];
        const float inv_cnt = 1.f / fmaxf(cnt, kF3ClampEps), invPm = 1.f / (float)P;
        for (int i = tid; i < D; i += 512) {
          p.pooled_l[(size_t)b * D + i] = (poolp[2 * D + i] + poolp[3 * D + i]) * inv_cnt;
          if (!pool_tc) p.pooled_v[(size_t)b * D + i] = (poolp[i] + poolp[D + i]) * invPm;
        }
      }
      f3_epi_bar();
    }
    stamp();

    // ---- E1: thread = patch.  (mb, column half) from the warp group; warps of a non-existent M block idle.
    const int mb = grp & 1, chh = grp >> 1;
    const int prow = 128 * mb + 32 * q + lane;          // patch of this thread
    const bool e1_act = mb < MB;
    const bool live = e1_act && prow < P;
    const float ivp = (e1_act && prow < NP) ? ivn[prow] : 0.f;
    const int cw = NT / 2, c_lo = chh * cw;             // NT / 2 is a multiple of 8
    const int combo = mb * 4 + q;
    const float invP = 1.f / (float)P;
    mbar_wait(s_full, 0);
    tc_fence_after();
    stamp();
    if (e1_act) {
      for (int c0 = c_lo; c0 < c_lo + cw; c0 += 8) {
        float x[8];
        tmem_ld8(tq + mb * NT + c0, x);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float s = x[j] * ivp * iln[c0 + j];
          const float lo = warp_redux_min(live ? s : CUDART_INF_F);
          const float hi = warp_redux_max(live ? s : -CUDART_INF_F);
          if (lane == j) *reinterpret_cast<float2*>(part_mm + (combo * NT + c0 + j) * 2) = make_float2(lo, hi);
        }
      }
    }
    stamp();
    f3_epi_bar();
    if (e1_act) {
      // every warp merges the partials of its own columns (identical arithmetic in every warp that shares a column)
      for (int j = lane; j < cw; j += 32) {
        float mn = CUDART_INF_F, mx = -CUDART_INF_F;
        for (int w = 0; w < 4 * MB; ++w) {
          const float2 t2 = *reinterpret_cast<const float2*>(part_mm + (w * NT + c_lo + j) * 2);
          mn = fminf(mn, t2.x); mx = fmaxf(mx, t2.y);
        }
        mnf[c_lo + j] = mn;
        irf[c_lo + j] = 1.f / (mx - mn + kF3MinMaxEps);
      }
      __syncwarp();
      for (int c0 = c_lo; c0 < c_lo + cw; c0 += 8) {
        float x[8], th[8], sr[8];
        tmem_ld8(tq + mb * NT + c0, x);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int t = c0 + j;
          const float s = x[j] * ivp * iln[t];
          const float mn = mnf[t];
          const float nn = (s - mn) * irf[t];
          const bool valid = msk[t] != 0.f;
          if (live && valid && s == mn) atomicMin(imn + t, prow);      // first arg-min patch, like torch.min
          th[j] = (live && valid && !(nn < p.thr)) ? nn : 0.f;
          if (pool_tc && t == T) th[j] = live ? invP : 0.f;           // spare column: G'[:, T] = mean_p v[p]
          sr[j] = live ? x[j] : 0.f;
        }
        if (prow < NP) {
          uint4 hi, lo;
          const uint32_t off = (uint32_t)((c0 >> 3) * NP + prow) * 16;
          split_hilo8(th, hi, lo);
          *reinterpret_cast<uint4*>(TH + off) = hi;
          *reinterpret_cast<uint4*>(TH + L.plane + off) = lo;
          split_hilo8(sr, hi, lo);
          *reinterpret_cast<uint4*>(SR + off) = hi;
          *reinterpret_cast<uint4*>(SR + L.plane + off) = lo;
        }
        const float cs = warp_colsum8(th, lane);
        if ((lane & 17) == 0) part_s[combo * NT + c0 + (lane >> 1)] = cs;
      }
    }
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(w_ready);
    stamp();
    f3_epi_bar();
    if (tid < NT) {
      float sg = 0.f;
      for (int w = 0; w < 4 * MB; ++w) sg += part_s[w * NT + tid];
      const bool valid = msk[tid] != 0.f;
      const float sigma = fmaxf(sg, kF3ClampEps);
      isg[tid] = valid ? 1.f / sigma : 0.f;
      if (tid < T) {
        float4 st4 = make_float4(mnf[tid], irf[tid], sigma, __int_as_float(imn[tid]));
        *reinterpret_cast<float4*>(p.stats + ((size_t)b * T + tid) * 4) = st4;
      }
    }
    f3_epi_bar();
    stamp();

    // ---- P1 epilogue: G'^T block [128 d][hi-part NT | lo-part NT] -> G = (hi + lo) / sigma_t: ||G_t||^2, bf16 hi | lo planes
    const int gw_ = NT / 4, g_lo = grp * gw_;           // NT / 4 is a multiple of 4
    const int dl = 32 * q + lane;                        // feature column inside the block
    float gn2[20];
#pragma unroll
    for (int k = 0; k < 20; ++k) gn2[k] = 0.f;
    for (int blk = 0; blk < L.NBLK; ++blk) {
      const int buf = blk & 1;
      mbar_wait(g_full + buf, (blk >> 1) & 1);
      tc_fence_after();
      const size_t dcol = (size_t)blk * 128 + dl;
      bf16* gh = p.g_split + ((size_t)b * 2) * T * D + dcol;
      bf16* gl = gh + (size_t)T * D;
      const uint32_t tg = tq + cG + buf * NT2 + g_lo;
#pragma unroll
      for (int c = 0; c < 20; c += 4) {                  // 4 columns at a time: hi-part and lo-part of the same tokens
        if (c < gw_) {
          float xh[4], xl[4];
          tmem_ld4(tg + c, xh);
          tmem_ld4(tg + NT + c, xl);
          tmem_ld_wait();
          if (c + 4 >= gw_) {                            // last chunk: the accumulator is in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(g_free + buf);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int t = g_lo + c + k;
            const float raw = xh[k] + xl[k];
            if (pool_tc && t == T) p.pooled_v[(size_t)b * D + dcol] = raw;
            const float g = raw * isg[t];
            gn2[c + k] = fmaf(g, g, gn2[c + k]);
            if (t < T) {
              const bf16 h = __float2bfloat16_rn(g);
              gh[(size_t)t * D] = h;
              gl[(size_t)t * D] = __float2bfloat16_rn(g - __bfloat162float(h));
            }
          }
        }
      }
      if (blk == 0) stamp();
    }
#pragma unroll
    for (int k = 0; k < 20; ++k) {
      if (k < gw_) {
        const float s = warp_sum(gn2[k]);
        if (lane == 0) part_s[q * NT + g_lo + k] = s;
      }
    }
    f3_epi_bar();
    if (tid < NT) {
      const float s = (part_s[tid] + part_s[NT + tid]) + (part_s[2 * NT + tid] + part_s[3 * NT + tid]);
      const float n = 1.f / fmaxf(sqrtf(s), kF3NormEps);
      ign[tid] = n;
      if (tid < T) p.g_inv_norm[(size_t)b * T + tid] = n;
    }
    f3_epi_bar();
    stamp();

    // ---- E3: logits[t][j] = s L'[t][j] / (sigma_t ||G_t|| ||l_j||)  (TMEM: lane = j, columns = t), masked LSE both ways
    mbar_wait(l_full, 0);
    tc_fence_after();
    float* Lb = reinterpret_cast<float*>(SR);           // [NT][NT + 1], the S_raw^T region is free once L' is done
    const int ldl = NT + 1;
    const int jrow = 32 * q + lane;
    if (32 * q < NT) {                                  // warp-uniform: tcgen05.ld needs the whole warp
      const bool jin = jrow < NT;
      const bool vj = jin && msk[jrow] != 0.f;
      const float ilj = jin ? iln[jrow] : 0.f;
#pragma unroll
      for (int c = 0; c < 20; c += 4) {
        if (c < gw_) {
          float xh[4], xl[4];
          tmem_ld4(tq + g_lo + c, xh);
          tmem_ld4(tq + NT + g_lo + c, xl);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int t = g_lo + c + k;
            const bool on = vj && msk[t] != 0.f;
            const float y = on ? (xh[k] + xl[k]) * (p.scale * isg[t] * ign[t]) * ilj : -CUDART_INF_F;
            if (jin) Lb[t * ldl + jrow] = y;
            if (p.tt_logits && t < T && jrow < T) p.tt_logits[((size_t)b * T + t) * T + jrow] = y;
          }
        }
      }
    }
    f3_epi_bar();
    stamp();
    float ce_r = 0.f, ce_c = 0.f;
    for (int t = ew; t < T; t += kF3EpiWarps) {          // row direction: softmax over j for token t (loss_vl_local)
      float m = -CUDART_INF_F;
      for (int j = lane; j < T; j += 32) m = fmaxf(m, Lb[t * ldl + j]);
      m = warp_max(m);
      float s = 0.f;
      for (int j = lane; j < T; j += 32) { const float y = Lb[t * ldl + j]; s += (y == -CUDART_INF_F) ? 0.f : __expf(y - m); }
      s = warp_sum(s);
      const bool valid = msk[t] != 0.f;
      const float lse = valid ? m + logf(s) : 0.f;
      if (lane == 0) { p.lse_row[(size_t)b * T + t] = lse; if (valid) ce_r += lse - Lb[t * ldl + t]; }
    }
    for (int j = ew; j < T; j += kF3EpiWarps) {          // column direction: softmax over t for token j (loss_lv_local)
      float m = -CUDART_INF_F;
      for (int t = lane; t < T; t += 32) m = fmaxf(m, Lb[t * ldl + j]);
      m = warp_max(m);
      float s = 0.f;
      for (int t = lane; t < T; t += 32) { const float y = Lb[t * ldl + j]; s += (y == -CUDART_INF_F) ? 0.f : __expf(y - m); }
      s = warp_sum(s);
      const bool valid = msk[j] != 0.f;
      const float lse = valid ? m + logf(s) : 0.f;
      if (lane == 0) { p.lse_col[(size_t)b * T + j] = lse; if (valid) ce_c += lse - Lb[j * ldl + j]; }
    }
    if (lane == 0) { red[ew] = ce_r; red[kF3EpiWarps + ew] = ce_c; }
    f3_epi_bar();
    if (tid == 0) {
      float a = 0.f, c = 0.f;
      for (int w = 0; w < kF3EpiWarps; ++w) { a += red[w]; c += red[kF3EpiWarps + w]; }
      p.local_partial[2 * b] = a;
      p.local_partial[2 * b + 1] = c;
    }
    stamp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

bool sparc_fwd3_supported(int P, int T, int D, int dtype) {
  if (dtype != CFA_DTYPE_BF16) return false;
  if (P < 1 || P > 256 || T < 1 || T > 80 || D % 128 || D < 128) return false;
  const Fwd3Layout L = fwd3_layout(P, T, D);
  return L.NS1 >= 2 && L.NS0 >= 2 && L.total + 1024 <= 227 * 1024;
}

int sparc_fwd3_launch(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, float thr,
                      float scale, float* row_inv_norm, float* pooled_v, float* pooled_l, float* lse_row, float* lse_col,
                      float* local_partial, float* tt_logits, float* g_inv_norm, void* g_split, float* stats,
                      long long* prof, int dtype, cudaStream_t st) {
  if (dtype != CFA_DTYPE_BF16) return CFA_ERR_UNSUPPORTED;
  if (!g_split || !stats) return CFA_ERR_WORKSPACE;
  const Fwd3Layout L = fwd3_layout(P, T, D);
  CUtensorMap tmV0, tmV1, tmL;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tmV0, v, D, P, B, 64, L.NP)) != CFA_OK) return rc;
  if ((rc = make_tmap_bf16_3d(&tmV1, v, D, P, B, 64, L.CR0)) != CFA_OK) return rc;
  if ((rc = make_tmap_bf16_3d(&tmL, l, D, T, B, 64, L.NT)) != CFA_OK) return rc;
  Fwd3Params prm{prof, P, T, D, thr, scale, mask, row_inv_norm, row_inv_norm + (size_t)B * P, pooled_v, pooled_l, lse_row,
                 lse_col, local_partial, tt_logits, g_inv_norm, (bf16*)g_split, stats};
  const size_t smem = L.total + 1024;
  CFA_CUDA_TRY(cudaFuncSetAttribute(sparc_fwd3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sparc_fwd3_kernel<false><<<B, kF3Threads, smem, st>>>(tmV0, tmV1, tmL, prm);
  return launch_status();
}

}  // namespace cfa
