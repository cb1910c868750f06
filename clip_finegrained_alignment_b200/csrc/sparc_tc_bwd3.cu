// Tensor-core SPARC backward, third generation ("transposed" orientation, see sparc_tc_fwd3.cu) for bf16 embeddings.
//
// Same mathematics as sparc_bwd2_kernel (SURVEY.md §8 a-bwd, gradients w.r.t. RAW dot products):
//
//   dG = dLhat . l - Gamma G             (Gamma = diag(gfac), the J_n term of normalize(G), losses.py:173)
//   dW = dG . v^T = dLhat . S_raw - Gamma Q ,  Q = G . v^T                   (losses.py:245 backward)
//   dv = (dShat + Z)^T . l + (-Gamma W)^T . G - v vfac + dvbar / P           Z = dLhat^T . W
//   dl = (dShat + Z) . v   - l lfac + m dlbar / cnt
//
// but every product with a raw tile has the raw tile as the A operand (M = patches or feature columns) and the on-chip
// bf16 hi|lo operand as B, stacked along N where that fits:
//
//   P1   [S^T | Qh^T | Ql^T][p, .] = v_kb . [l_kb ; Ghi_kb ; Glo_kb]^T       ONE N = 240 MMA per k-step: S_raw and Q = G . v^T
//        phase 0 (overlaps P1): saved T x T logits -> dLhat (hi|lo operand), gfac_t, column sums
//   E1   thread = patch: S^T -> W^T, S_raw^T (hi|lo operands [t/8][p][t%8]); -gfac Q parked in TMEM (tcgen05.st)
//   M    dW^T += S_raw^T . dLhat^T (on top of -gfac Q) ;  Z^T = W^T . dLhat
//   E3   renorm / threshold / min-max backward (one sweep, identities of sparc_bwd2) -> dShat'^T (hi|lo), -gfac W^T (hi|lo)
//   P4   per 128-wide D block:  dl^T[d,t] = v^T . dShat'^T   (N = 160, hi|lo stacked)
//                               dv^T[d,p] = l^T . dShat' + G^T . (-gfac W)      (N = NP; hi, lo operands in turn)
//        epilogue: - x fac, pooled-mean terms, bf16 -> global; thread = feature column, 64 contiguous bytes per warp store
//
// Template parameters kNT / kNP / kD as in the forward (0 = run-time values).
#include "tc_common.cuh"
#include "sparc_paths.h"
#include <math_constants.h>

namespace cfa {
using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int kB3EpiWarps = 16;
constexpr int kB3Threads = 32 * (2 + kB3EpiWarps);   // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue
constexpr float kB3ClampEps = 1e-8f;

struct Bwd3Layout {
  int NP, NT, MB, KB0, NBLK, CR0, NCH, NSP;
  uint32_t v_bytes, l_bytes, slotP, slot4, plane, dlb;
  uint32_t off_sr, off_w, off_ring4, off_stage, off_ldp, off_dl, off_f, off_bar, total;
};

__host__ __device__ inline Bwd3Layout bwd3_layout(int P, int T, int D) {
  Bwd3Layout L;
  L.NP = (P + 15) & ~15; L.NT = (T + 15) & ~15; L.MB = L.NP > 128 ? 2 : 1; L.KB0 = D / 64; L.NBLK = D / 128;
  // P4 streams v in chunks of NT patch rows, so that every P4 tile pair (v chunk, l, G hi, G lo) is [NT x 128 d]
  L.CR0 = L.NT;
  L.NCH = (L.NP + L.NT - 1) / L.NT;
  L.v_bytes = (uint32_t)L.NP * 128; L.l_bytes = (uint32_t)L.NT * 128;
  L.slotP = L.v_bytes + 3 * L.l_bytes;
  L.slot4 = 2u * L.l_bytes;
  L.plane = (uint32_t)L.NT * L.NP * 2;
  L.dlb = (uint32_t)L.NT * L.NT * 2;
  const uint32_t op = (2 * L.plane + 1023) & ~1023u;
  L.off_sr = 0; L.off_w = op; L.off_ring4 = 2 * op;
  const uint32_t ldp = (uint32_t)kB3EpiWarps * L.NT * 4;
  const uint32_t nf = 2u * L.NP + 18u * L.NT + 64;
  const uint32_t budget = 227u * 1024u - 1024u;
  L.NSP = 3;
  for (;;) {
    uint32_t p1_end = L.NSP * L.slotP;
    const uint32_t reach = (L.NSP - 1) * L.slotP + 256u * 128u;          // M block 1 of the last P1 slot
    if (reach > p1_end) p1_end = reach;
    uint32_t dl0 = L.off_ring4 + 2 * L.slot4;
    if (p1_end + ldp > dl0) dl0 = p1_end + ldp;
    dl0 = (dl0 + 127) & ~127u;
    uint32_t end = dl0 + 2 * L.dlb;
    // per-warp transposition tiles of the P4 output epilogue [8][36] fp32, behind the P4 ring (used after ds_ready only)
    L.off_stage = L.off_ring4 + 3 * L.slot4;
    if (L.off_stage + (uint32_t)kB3EpiWarps * 8u * 36u * 4u > end) end = L.off_stage + (uint32_t)kB3EpiWarps * 8u * 36u * 4u;
    // scratch of E3' (column partials [2][8][NT], row partials [2][NP]) lives in the dead dLhat region
    const uint32_t e3 = (16u * L.NT + 2u * L.NP) * 4;
    if (dl0 + e3 > end) end = dl0 + e3;
    if (L.off_ring4 + 3 * L.slot4 > end) end = L.off_ring4 + 3 * L.slot4;
    // M block 1 of the K-major views of the interleaved operands reads up to row 255 of the last chunk
    if (L.off_w + 2 * L.plane + 256u * 16u > end) end = L.off_w + 2 * L.plane + 256u * 16u;
    L.off_dl = dl0; L.off_ldp = dl0 - ldp;
    L.off_f = (end + 127) & ~127u;
    L.off_bar = (L.off_f + 4 * nf + 7) & ~7u;
    L.total = L.off_bar + 8 * 32;
    if (L.total <= budget || L.NSP == 2) break;
    L.NSP = 2;
  }
  return L;
}

struct Bwd3Params {
  long long* prof;
  int P, T, D;
  float thr, scale;
  const uint8_t* mask;
  const float* inv_vn;
  const float* inv_ln;
  const float* lse_row;
  const float* lse_col;
  const float* coef;
  const float* tt_logits;   // [B][T][T] masked, scaled logits from the forward
  const float* g_inv_norm;  // [B][T]
  const float* stats;       // [B][T][4] min, 1/range, sigma, arg-min patch
  const float* dpool_v;
  const float* dpool_l;
  CoefSrc cs;               // cs.on: coefficients from the upstream-gradient pointers instead of `coef`
  int pdl_late;             // != 0: griddepcontrol.wait right before the first read of dpool_*, else at kernel start
  const bf16* v;
  const bf16* l;
  bf16* dv;
  bf16* dl;
};

__device__ __forceinline__ void b3_epi_bar() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ float b3_raw(const bf16* p) {
  return __uint_as_float((uint32_t)(*reinterpret_cast<const unsigned short*>(p)) << 16);
}

template <int kNT, int kNP, int kD, bool kHalf>
__global__ void __launch_bounds__(kB3Threads, 1)
sparc_bwd3_kernel(const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmV1,
                  const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmG, const Bwd3Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if (!p.pdl_late) pdl_wait();                        // ordinary stream order (see g_sparc_bwd_pdl_late)
  uint8_t* base = CFA_SMEM_BASE_1024(smem_raw);
  const Bwd3Layout L = bwd3_layout(p.P, p.T, kD ? kD : p.D);
  const int NP = kNP ? kNP : L.NP, NT = kNT ? kNT : L.NT, D = kD ? kD : p.D;
  const int MB = NP > 128 ? 2 : 1, KB0 = D / 64, NBLK = D / 128;
  const int CR0 = NT, NCH = (NP + NT - 1) / NT;       // P4: v in chunks of NT patch rows
  const int NSP = L.NSP, P = p.P, T = p.T;
  const int NT2 = 2 * NT, NT3 = 3 * NT;
  const uint32_t v_bytes = (uint32_t)NP * 128, l_bytes = (uint32_t)NT * 128, slotP = v_bytes + 3 * l_bytes;
  const uint32_t slot4 = 2u * l_bytes;
  const uint32_t plane = (uint32_t)NT * NP * 2, dlb = (uint32_t)NT * NT * 2;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* SR = base + L.off_sr;                      // S_raw^T hi|lo, later dShat'^T hi|lo : [2 NT / 8][NP][8]
  uint8_t* WT = base + L.off_w;                       // W^T hi|lo, later (-gfac W)^T hi|lo
  uint8_t* ringP = base;                              // P1 slots [v tile | l tile | G hi tile | G lo tile]
  uint8_t* ring4 = base + L.off_ring4;                // P4 slots: v chunk pair, or l / G hi / G lo tile pair
  uint8_t* DLh = base + L.off_dl;                     // dLhat hi, lo : [NT / 8][NT][8] each
  uint8_t* DLl = DLh + dlb;
  float* ldpart = (float*)(base + L.off_ldp);         // [16][NT] phase-0 column partials
  float* part_c = (float*)(base + L.off_dl);          // E3' scratch in the dead dLhat region: [8][NT] sum of kept dW
  float* part_d = part_c + 8 * NT;                    //                                       [8][NT] sum of ds * s
  float* vqp = part_d + 8 * NT;                       //                                       [2][NP] row partials
  float* ivn = (float*)(base + L.off_f);              // [NP]
  float* vfac = ivn + NP;                             // [NP]
  float4* cA = (float4*)(vfac + NP);                  // [NT] {1/||l_t||, min (+inf: masked), 1/range, 1/(sigma range)}
  float2* cD = (float2*)(cA + NT);                    // [NT] {1/sigma (0: masked), -gfac}
  float* msk = (float*)(cD + NT);                     // [NT] ...
  float* lser = msk + NT;
  float* lsec = lser + NT;
  float* mnf = lsec + NT;
  float* ignv = mnf + NT;
  float* ldot = ignv + NT;
  float* lfacs = ldot + NT;
  float* dmns = lfacs + NT;
  float* fixv = dmns + NT;
  float* mdl = fixv + NT;                             // msk / cnt
  int* imn = (int*)(mdl + NT);
  uint64_t* bars = (uint64_t*)(base + L.off_bar);
  uint64_t* fullP = bars;           // [3]
  uint64_t* emptyP = bars + 3;      // [3]
  uint64_t* full4 = bars + 6;       // [3]
  uint64_t* empty4 = bars + 9;      // [3]
  uint64_t* s_full = bars + 12;
  uint64_t* dl_ready = bars + 13;
  uint64_t* e1_ready = bars + 14;
  uint64_t* dw_full = bars + 15;
  uint64_t* ds_ready = bars + 16;
  uint64_t* oa_full = bars + 17;
  uint64_t* oa_free = bars + 18;
  uint64_t* ob_full = bars + 19;
  uint64_t* ob_free = bars + 20;
  uint32_t* tmem_slot = (uint32_t*)(bars + 22);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(fullP + i, 1); mbar_init(emptyP + i, 1); mbar_init(full4 + i, 1); mbar_init(empty4 + i, 1); }
    mbar_init(s_full, 1); mbar_init(dl_ready, kB3EpiWarps); mbar_init(e1_ready, kB3EpiWarps); mbar_init(dw_full, 1);
    mbar_init(ds_ready, kB3EpiWarps);
    mbar_init(oa_full, 1); mbar_init(oa_free, kB3EpiWarps); mbar_init(ob_full, 1); mbar_init(ob_free, kB3EpiWarps);
    fence_barrier_init();
    tma_prefetch_desc(&tmV0); tma_prefetch_desc(&tmV1); tma_prefetch_desc(&tmL); tma_prefetch_desc(&tmG);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < NP + NT; i += kB3Threads) {
    if (i < NP) ivn[i] = (i < P) ? p.inv_vn[(size_t)b * P + i] : 0.f;
    else {
      const int t = i - NP;
      const bool in = t < T;
      const bool on = in && p.mask[(size_t)b * T + t];
      msk[t] = on ? 1.f : 0.f;
      lser[t] = in ? p.lse_row[(size_t)b * T + t] : 0.f;
      lsec[t] = in ? p.lse_col[(size_t)b * T + t] : 0.f;
      ignv[t] = in ? p.g_inv_norm[(size_t)b * T + t] : 0.f;
      float4 st4 = make_float4(0.f, 1.f, 1.f, __int_as_float(0x7fffffff));      // finite 1/range: (s - inf) * 0 would be NaN
      if (in) st4 = __ldg(reinterpret_cast<const float4*>(p.stats + ((size_t)b * T + t) * 4));
      const float isg = on ? 1.f / st4.z : 0.f;
      mnf[t] = st4.x;
      imn[t] = __float_as_int(st4.w);
      cA[t] = make_float4(in ? p.inv_ln[(size_t)b * T + t] : 0.f, on ? st4.x : CUDART_INF_F, st4.y, isg * st4.y);
      cD[t].x = isg;
    }
  }
  if (warp == 2) {                                     // mdl[t] = m_t / (number of valid tokens)   (losses.py:211 backward)
    float c = 0.f;
    for (int t = lane; t < T; t += 32) c += p.mask[(size_t)b * T + t] ? 1.f : 0.f;
    c = 1.f / fmaxf(warp_sum(c), kB3ClampEps);
    for (int t = lane; t < NT; t += 32) mdl[t] = (t < T && p.mask[(size_t)b * T + t]) ? c : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t cDL = 0, cDV = (uint32_t)NT2;        // P4 accumulators

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      for (int u = 0; u < KB0; ++u) {                  // P1: v | l | G hi | G lo
        const int s = u % NSP;
        if (u >= NSP) mbar_wait_sleep(emptyP + s, ((u / NSP) - 1) & 1);
        uint8_t* st = ringP + (size_t)s * slotP;
        mbar_expect_tx(fullP + s, slotP);
        tma_load_3d(st, &tmV0, fullP + s, u * 64, 0, b);
        tma_load_3d(st + v_bytes, &tmL, fullP + s, u * 64, 0, b);
        tma_load_3d(st + v_bytes + l_bytes, &tmG, fullP + s, u * 64, 0, 2 * b);
        tma_load_3d(st + v_bytes + 2 * l_bytes, &tmG, fullP + s, u * 64, 0, 2 * b + 1);
      }
      // P4 slots 0, 1 overlap the P1 ring and the phase-0 scratch; slot 2 overlaps dLhat and the E3' scratch
      mbar_wait_sleep(s_full, 0);
      mbar_wait_sleep(dl_ready, 0);
      const int per = NCH + 3, n4 = NBLK * per;
      for (int i = 0; i < n4; ++i) {
        const int s = i % 3, blk = i / per, w = i % per;
        if (i >= 3) mbar_wait_sleep(empty4 + s, ((i / 3) - 1) & 1);
        if (i == 2) mbar_wait_sleep(ds_ready, 0);
        uint8_t* st = ring4 + (size_t)s * slot4;
        if (w < NCH) {
          mbar_expect_tx(full4 + s, 2 * l_bytes);
          tma_load_3d(st, &tmV1, full4 + s, blk * 128, w * CR0, b);
          tma_load_3d(st + l_bytes, &tmV1, full4 + s, blk * 128 + 64, w * CR0, b);
        } else {
          const CUtensorMap* tm = (w == NCH) ? &tmL : &tmG;
          const int pl = (w == NCH) ? b : (w == NCH + 1 ? 2 * b : 2 * b + 1);
          mbar_expect_tx(full4 + s, 2 * l_bytes);
          tma_load_3d(st, tm, full4 + s, blk * 128, 0, pl);
          tma_load_3d(st + l_bytes, tm, full4 + s, blk * 128 + 64, 0, pl);
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (warp-uniform control flow, one elected lane issues) ===============================
    const bool leader = elect_one();
    const uint32_t id_p1 = make_idesc16(128, NT3, false, false, kHalf, kHalf);      // raw v x [l ; G hi ; G lo], all K-major
    const uint32_t id_dw = make_idesc16(128, NT, false, false, kHalf, kHalf);      // S_raw^T (K-major) x dLhat (K-major: N = t)
    const uint32_t id_z = make_idesc16(128, NT, false, true, kHalf, kHalf);        // W^T (K-major) x dLhat (MN-major: N = j)
    const uint32_t id_dl = make_idesc16(128, NT2, true, true, kHalf, kHalf);       // raw v^T (MN-major) x dShat'^T hi|lo (MN-major)
    const uint32_t id_dv = make_idesc16(128, NP, true, false, kHalf, kHalf);       // raw l^T (MN-major) x dShat' (K-major: N = p)
    const uint32_t id_dg = make_idesc16(128, NP, true, false, kHalf, kHalf);       // G^T (MN-major, bf16) x -gfac W (K-major)
    const uint64_t sw0 = make_smem_desc(0, 16, 1024, kLayoutSw128);
    long long* pf = (p.prof && leader) ? p.prof + (size_t)b * 32 : nullptr;
    int pi = 0;
    auto stamp = [&]() { if (pf) pf[pi++] = clock64(); };
    stamp();
    // ---- P1
    for (int u = 0, s = 0, ph = 0; u < KB0; ++u) {
      mbar_wait_sleep(fullP + s, ph);
      tc_fence_after();
      const uint32_t sv = smem_u32(ringP + (size_t)s * slotP), sb = sv + v_bytes;
      const uint64_t dv0 = sw0 | (sv >> 4), db0 = sw0 | (sb >> 4);
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        if (mb < MB) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss_w(leader, tmem + mb * NT3, dv0 + mb * (16384 >> 4) + 2 * k, db0 + 2 * k, id_p1, (u | k) != 0);
        }
      }
      umma_commit_w(leader, emptyP + s);
      if (++s == NSP) { s = 0; ph ^= 1; }
    }
    umma_commit_w(leader, s_full);
    stamp();
    // ---- M: dW^T (on top of the parked -gfac Q) and Z^T
    mbar_wait_sleep(dl_ready, 0);
    mbar_wait_sleep(e1_ready, 0);
    tc_fence_after();
    stamp();
    const uint32_t np16 = (uint32_t)NP * 16, nt16 = (uint32_t)NT * 16;
    // interleaved [k/8][rows][8] operands: K-major: LBO = rows * 16, SBO = 128; MN-major: LBO = 128, SBO = rows * 16
    const uint64_t k_srh = make_smem_desc(smem_u32(SR), np16, 128, kLayoutNone), k_srl = make_smem_desc(smem_u32(SR) + plane, np16, 128, kLayoutNone);
    const uint64_t k_wh = make_smem_desc(smem_u32(WT), np16, 128, kLayoutNone), k_wl = make_smem_desc(smem_u32(WT) + plane, np16, 128, kLayoutNone);
    const uint64_t k_dlh = make_smem_desc(smem_u32(DLh), nt16, 128, kLayoutNone), k_dll = make_smem_desc(smem_u32(DLl), nt16, 128, kLayoutNone);
    const uint64_t m_dlh = make_smem_desc(smem_u32(DLh), 128, nt16, kLayoutNone), m_dll = make_smem_desc(smem_u32(DLl), 128, nt16, kLayoutNone);
    const uint32_t ksA = (2 * np16) >> 4;              // K-major interleaved with NP rows: two 8-wide chunks per k-step
    const uint32_t ksB = (2 * nt16) >> 4;              // K-major interleaved with NT rows
    const uint32_t mtile = (128u * 16u) >> 4;          // second 128-row M block of a K-major interleaved operand
    const int nksT = NT / 16;
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
      if (mb < MB) {
        const uint32_t dw = tmem + mb * NT3 + NT, dz = tmem + mb * NT3 + NT2;
        const uint32_t mo = mb * mtile;
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) if (ks < nksT) umma_ss_w(leader, dw, k_srh + mo + ks * ksA, k_dlh + ks * ksB, id_dw, true);
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) if (ks < nksT) umma_ss_w(leader, dw, k_srh + mo + ks * ksA, k_dll + ks * ksB, id_dw, true);
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) if (ks < nksT) umma_ss_w(leader, dw, k_srl + mo + ks * ksA, k_dlh + ks * ksB, id_dw, true);
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) if (ks < nksT) umma_ss_w(leader, dz, k_wh + mo + ks * ksA, m_dlh + ks * 16, id_z, ks != 0);
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) if (ks < nksT) umma_ss_w(leader, dz, k_wh + mo + ks * ksA, m_dll + ks * 16, id_z, true);
#pragma unroll
        for (int ks = 0; ks < 5; ++ks) if (ks < nksT) umma_ss_w(leader, dz, k_wl + mo + ks * ksA, m_dlh + ks * 16, id_z, true);
      }
    }
    umma_commit_w(leader, dw_full);
    // ---- P4
    mbar_wait_sleep(ds_ready, 0);
    tc_fence_after();
    stamp();
    const uint64_t m_ds = make_smem_desc(smem_u32(SR), 128, np16, kLayoutNone);          // dShat'^T hi|lo, MN-major (N = t, K = p)
    int s4 = 0, ph4 = 0;
    long long wfull = 0, wfree = 0;
    for (int blk = 0; blk < NBLK; ++blk) {
      // unit A: dl^T
      { const long long w0 = clock64(); mbar_wait_sleep(oa_free, (blk & 1) ^ 1); wfree += clock64() - w0; }
      tc_fence_after();
      for (int ch = 0; ch < NCH; ++ch) {
        {
          const int s = s4;
          { const long long w0 = clock64(); mbar_wait_sleep(full4 + s, ph4); wfull += clock64() - w0; }
          tc_fence_after();
          const uint32_t sa = smem_u32(ring4 + (size_t)s * slot4);
          const uint64_t da = make_smem_desc(sa, l_bytes, 1024, kLayoutSw128);
          const int r0 = ch * CR0, nk = min(CR0, NP - r0) / 16;
          const uint64_t db = m_ds + (uint32_t)r0;
#pragma unroll
          for (int ks = 0; ks < 5; ++ks) if (ks < nk) umma_ss_w(leader, tmem + cDL, da + ks * 128, db + ks * 16, id_dl, (ch | ks) != 0);
          umma_commit_w(leader, empty4 + s);
          if (++s4 == 3) { s4 = 0; ph4 ^= 1; }
        }
      }
      umma_commit_w(leader, oa_full);
      // unit B: dv^T
      { const long long w0 = clock64(); mbar_wait_sleep(ob_free, (blk & 1) ^ 1); wfree += clock64() - w0; }
      tc_fence_after();
#pragma unroll
      for (int w = 0; w < 3; ++w) {
        const int s = s4;
        { const long long w0 = clock64(); mbar_wait_sleep(full4 + s, ph4); wfull += clock64() - w0; }
        tc_fence_after();
        const uint32_t sa = smem_u32(ring4 + (size_t)s * slot4);
        const uint64_t da = make_smem_desc(sa, l_bytes, 1024, kLayoutSw128);             // [NT x 64] tile pair, MN-major (M = d)
        if (w == 0) {
#pragma unroll
          for (int ks = 0; ks < 5; ++ks) if (ks < nksT) umma_ss_w(leader, tmem + cDV, da + ks * 128, k_srh + ks * ksA, id_dv, ks != 0);
#pragma unroll
          for (int ks = 0; ks < 5; ++ks) if (ks < nksT) umma_ss_w(leader, tmem + cDV, da + ks * 128, k_srl + ks * ksA, id_dv, true);
        } else if (w == 1) {
#pragma unroll
          for (int ks = 0; ks < 5; ++ks) if (ks < nksT) umma_ss_w(leader, tmem + cDV, da + ks * 128, k_wh + ks * ksA, id_dg, true);
#pragma unroll
          for (int ks = 0; ks < 5; ++ks) if (ks < nksT) umma_ss_w(leader, tmem + cDV, da + ks * 128, k_wl + ks * ksA, id_dg, true);
        } else {
#pragma unroll
          for (int ks = 0; ks < 5; ++ks) if (ks < nksT) umma_ss_w(leader, tmem + cDV, da + ks * 128, k_wh + ks * ksA, id_dg, true);
        }
        umma_commit_w(leader, empty4 + s);
        if (++s4 == 3) { s4 = 0; ph4 ^= 1; }
      }
      umma_commit_w(leader, ob_full);
    }
    stamp();
    if (pf) { pf[8] = wfull; pf[9] = wfree; }
  } else {
    // =============================== epilogue: 16 warps = 4 TMEM lane quarters x 4 groups ===============================
    const int ew = warp - 2, q = warp & 3, grp = ew >> 2;
    const uint32_t tq = tmem + ((uint32_t)(32 * q) << 16);
    const int tid = ew * 32 + lane;
    float c_r, c_c;
    if (p.cs.on) {
      float c4[4];
      coef_from_src(p.cs, c4);
      c_r = c4[2]; c_c = c4[3];
    } else {
      c_r = p.coef[0]; c_c = p.coef[1];
    }
    long long* pf = (p.prof && tid == 0) ? p.prof + (size_t)b * 32 + 16 : nullptr;
    int pi = 0;
    const bool pwarp = p.prof != nullptr && ew == 0;    // warp-uniform: the other 15 warps skip a stamp with one branch
    auto stamp = [&]() { if (pwarp) { if (pf) pf[pi] = clock64(); ++pi; } };
    stamp();
    // fp16 operands (kind::f16 takes no mixed fp16 x bf16 pair, so the on-chip operands are fp16 hi|lo): two per-sample powers
    // of two keep them in fp16's normal range.  sa <= 1 / (max ||v|| max ||l||) scales S_raw (the forward's value, recomputed
    // from the stored norms); sc ~ 1 / max |dLhat| scales the gradient chain, which is linear in it -- a 2^16 loss scale
    // (GradScaler) or a 1e-6 coefficient both land at O(1).  The outputs are multiplied by 1 / sc (exact).
    float sa = 1.f, isa = 1.f, sc = 1.f, isc = 1.f;
    if (kHalf) {
      float mv = CUDART_INF_F, ml = CUDART_INF_F, xg = 0.f, xl = 0.f;
      for (int i = lane; i < NP; i += 32) { const float n = ivn[i]; mv = (n > 0.f) ? fminf(mv, n) : mv; }
      for (int i = lane; i < NT; i += 32) {
        const float n = cA[i].x;
        ml = (n > 0.f) ? fminf(ml, n) : ml;
        xl = fmaxf(xl, n);
        xg = fmaxf(xg, (msk[i] != 0.f) ? ignv[i] : 0.f);
      }
      mv = warp_redux_min(mv); ml = warp_redux_min(ml); xg = warp_redux_max(xg); xl = warp_redux_max(xl);
      sa = pow2_floor_clamped(mv * ml);
      isa = 1.f / sa;
      const float bound = fabsf(p.scale) * (fabsf(c_r) + fabsf(c_c)) * xg * xl;
      sc = pow2_floor_clamped(1.f / fmaxf(bound, 1e-37f));
      isc = 1.f / sc;
    }

    // ---- phase 0 (overlaps P1): saved T x T logits -> dLhat (hi|lo operand [j/8][t][j%8]), gfac_t, ldot_j.
    // A warp covers rpp rows at a time: lane -> (row slot rs, 8-column chunk jc); fixed-order sums (deterministic bits).
    {
      const int nch = NT / 8, rpp = 32 / nch;
      const int rs = lane / nch, jc = lane - rs * nch;
      const bool lact = rs < rpp;
      float cs[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) cs[k] = 0.f;
      for (int t0 = ew * rpp; t0 < NT; t0 += kB3EpiWarps * rpp) {
        const int t = t0 + rs;
        const bool tact = lact && t < NT;
        const bool vt = tact && t < T && msk[t] != 0.f;
        float y[8], x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = 0.f;
        if (vt) {
          const float* src = p.tt_logits + ((size_t)b * T + t) * T + 8 * jc;
#pragma unroll
          for (int k = 0; k < 8; ++k) if (8 * jc + k < T) y[k] = __ldg(src + k);
        }
        const float lr = tact ? lser[t] : 0.f, ig = tact ? ignv[t] : 0.f;
        float gd = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int j = 8 * jc + k;
          const bool on = vt && j < T && msk[j] != 0.f;
          const float yy = on ? y[k] : 0.f;
          float g = c_r * __expf(fminf(yy - lr, 0.f)) + c_c * __expf(fminf(yy - lsec[j], 0.f));
          g -= (j == t) ? (c_r + c_c) : 0.f;
          g = on ? g : 0.f;
          const float pr = g * yy;
          gd += pr;
          cs[k] += pr;
          x[k] = p.scale * g * ig * cA[j].x * sc;
        }
        if (tact) {
          uint4 hi, lo;
          split_hilo8_t<kHalf>(x, hi, lo);
          const uint32_t off = (uint32_t)(jc * NT + t) * 16;
          *reinterpret_cast<uint4*>(DLh + off) = hi;
          *reinterpret_cast<uint4*>(DLl + off) = lo;
        }
        // row sum over the nch chunk lanes of this row slot (fixed order)
        float tot = 0.f;
        for (int k = 0; k < nch; ++k) tot += __shfl_sync(0xffffffffu, gd, (rs < rpp ? rs : 0) * nch + k);
        if (tact && jc == 0) cD[t].y = -(tot * ig * ig);         // -gfac_t = -(g^_t . dg^_t) / ||G_t||^2
      }
      // column partials of this warp: add the rpp row slots in fixed order, then one slot per warp
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float tot = 0.f;
        for (int r = 0; r < rpp; ++r) tot += __shfl_sync(0xffffffffu, cs[k], r * nch + (lane < nch ? lane : 0));
        if (lane < nch) ldpart[ew * NT + 8 * lane + k] = tot;
      }
      b3_epi_bar();
      if (tid < NT) {
        float s = 0.f;
        for (int w = 0; w < kB3EpiWarps; ++w) s += ldpart[w * NT + tid];
        ldot[tid] = s * sc;
      }
      fence_proxy_async();
      b3_epi_bar();
      if (lane == 0) mbar_arrive(dl_ready);
    }
    stamp();

    // ---- E1: thread = patch.  S^T -> W^T (hi|lo), S_raw^T (hi|lo); -gfac Q parked in the dW^T columns
    const int mb = grp & 1, chh = grp >> 1;
    const int prow = 128 * mb + 32 * q + lane;
    const bool e_act = mb < MB;
    const bool live = e_act && prow < P;
    const float ivp = (e_act && prow < NP) ? ivn[prow] : 0.f;
    const int cw = NT / 2, c_lo = chh * cw;
    const int combo = mb * 4 + q;
    const uint32_t tS = tq + mb * NT3, tW = tS + NT, tZ = tS + NT2;
    mbar_wait_sleep(s_full, 0);
    tc_fence_after();
    stamp();
    if (e_act) {
      // software-pipelined TMEM reads (here and in E3'): the loads of chunk g8 + 1 are issued once chunk g8's inputs are
      // consumed, so their latency overlaps the hi|lo splits and the shared-memory / TMEM stores of chunk g8
      float x[8], qh[8], ql[8];
      tmem_ld8(tS + c_lo, x);
      tmem_ld8(tW + c_lo, qh);
      tmem_ld8(tZ + c_lo, ql);
#pragma unroll 1                                   // one hot loop body instead of 5 cold copies (instruction fetch)
      for (int g8 = 0; g8 < 5; ++g8) {
        const int c0 = c_lo + 8 * g8;
        if (8 * g8 < cw) {
          float w[8], sr[8], qq[8];
          tmem_ld_wait8(x);
          tmem_ld_wait8(qh);
          tmem_ld_wait8(ql);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 a4 = cA[c0 + j];
            const float2 d2 = cD[c0 + j];
            const float raw = live ? x[j] : 0.f;
            const float nn = __fmul_rn(__fsub_rn(__fmul_rn(__fmul_rn(raw, ivp), a4.x), a4.y), a4.z);   // the forward's roundings
            w[j] = (live && !(nn < p.thr)) ? nn * d2.x : 0.f;
            sr[j] = kHalf ? raw * sa : raw;
            qq[j] = (kHalf ? d2.y * (sa * sc) : d2.y) * (qh[j] + ql[j]);     // -gfac Q (gfac = 0 for masked tokens), fp16: x sa sc like the dW sum
          }
          if (8 * g8 + 8 < cw) {
            tmem_ld8(tS + c0 + 8, x);
            tmem_ld8(tW + c0 + 8, qh);
            tmem_ld8(tZ + c0 + 8, ql);
          }
          if (prow < NP) {
            uint4 hi, lo;
            const uint32_t off = (uint32_t)((c0 >> 3) * NP + prow) * 16;
            split_hilo8_t<kHalf>(w, hi, lo);
            *reinterpret_cast<uint4*>(WT + off) = hi;
            *reinterpret_cast<uint4*>(WT + plane + off) = lo;
            split_hilo8_t<kHalf>(sr, hi, lo);
            *reinterpret_cast<uint4*>(SR + off) = hi;
            *reinterpret_cast<uint4*>(SR + plane + off) = lo;
          }
          tmem_st8(tW + c0, qq);
        }
      }
      tmem_st_wait();
    }
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(e1_ready);
    stamp();

    // ---- E3': dW -> renorm / threshold / min-max backward -> dShat'^T = dShat^T + Z^T (hi|lo), (-gfac W)^T (hi|lo), lfac, vfac
    // One sweep, with the two identities of sparc_bwd2:  sum_p W dW = 0  and  sum_p dN N = 0  (only the arg-min patch
    // receives a scatter term, patched in afterwards).
    mbar_wait_sleep(dw_full, 0);
    tc_fence_after();
    stamp();
    float vq = 0.f;
    if (e_act) {
      float x[8], dw[8], z[8];
      tmem_ld8(tS + c_lo, x);
      tmem_ld8(tW + c_lo, dw);
      tmem_ld8(tZ + c_lo, z);
#pragma unroll 1                                   // one hot loop body instead of 5 cold copies (instruction fetch)
      for (int g8 = 0; g8 < 5; ++g8) {
        const int c0 = c_lo + 8 * g8;
        if (8 * g8 < cw) {
          float dsh[8], dk[8], pr[8], wg[8];
          tmem_ld_wait8(x);
          tmem_ld_wait8(dw);
          tmem_ld_wait8(z);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 a4 = cA[c0 + j];
            const float2 d2 = cD[c0 + j];
            const float u = live ? __fmul_rn(x[j], ivp) : 0.f;      // phantom lanes may hold NaN/Inf: select, never 0 * x
            const float s = __fmul_rn(u, a4.x);
            const float nn = __fmul_rn(__fsub_rn(s, a4.y), a4.z);
            const bool kept = live && !(nn < p.thr);
            const float d = kept ? (kHalf ? dw[j] * isa : dw[j]) : 0.f;          // fp16: sc dW from here on
            dk[j] = d;
            const float ds = d * a4.w;                    // / (sigma range)
            const float prod = ds * s;
            pr[j] = prod;
            vq += prod;
            dsh[j] = fmaf(ds * a4.x, ivp, live ? z[j] : 0.f);           // dShat' = dShat + Z
            wg[j] = kept ? (nn * d2.x) * (kHalf ? d2.y * sc : d2.y) : 0.f;       // -gfac W
          }
          if (prow < NP) {
            uint4 hi, lo;
            const uint32_t off = (uint32_t)((c0 >> 3) * NP + prow) * 16;
            split_hilo8_t<kHalf>(dsh, hi, lo);
            *reinterpret_cast<uint4*>(SR + off) = hi;
            *reinterpret_cast<uint4*>(SR + plane + off) = lo;
            split_hilo8_t<kHalf>(wg, hi, lo);
            *reinterpret_cast<uint4*>(WT + off) = hi;
            *reinterpret_cast<uint4*>(WT + plane + off) = lo;
          }
          if (8 * g8 + 8 < cw) {                          // issued after the splits (register pressure); overlaps the column sums
            tmem_ld8(tS + c0 + 8, x);
            tmem_ld8(tW + c0 + 8, dw);
            tmem_ld8(tZ + c0 + 8, z);
          }
          const float c1 = warp_colsum8(dk, lane);
          const float c2 = warp_colsum8(pr, lane);
          if ((lane & 17) == 0) { part_c[combo * NT + c0 + (lane >> 1)] = c1; part_d[combo * NT + c0 + (lane >> 1)] = c2; }
        }
      }
    }
    b3_epi_bar();
    if (tid < NT) {
      float cx = 0.f, sd = 0.f;
      for (int w = 0; w < 4 * MB; ++w) { cx += part_c[w * NT + tid]; sd += part_d[w * NT + tid]; }
      const float4 a4 = cA[tid];
      const bool valid = msk[tid] != 0.f;
      const float dmn = valid ? -cx * a4.w : 0.f;
      dmns[tid] = dmn;
      fixv[tid] = dmn * mnf[tid];                       // ds[imn] * s[imn], s[imn] = min
      lfacs[tid] = (sd + dmn * mnf[tid] + ldot[tid]) * a4.x * a4.x;      // (l^_t . dl^_t) / ||l_t||^2
    }
    b3_epi_bar();
    if (e_act) {
      for (int t = c_lo; t < c_lo + cw; ++t) {
        if (imn[t] == prow && msk[t] != 0.f && live) {  // this thread owns the arg-min element of token t
          const uint32_t off = (uint32_t)((t >> 3) * NP + prow) * 16 + (t & 7) * 2;
          if (kHalf) {
            __half* ph = reinterpret_cast<__half*>(SR + off);
            __half* pl = reinterpret_cast<__half*>(SR + plane + off);
            const float f = (__half2float(*ph) + __half2float(*pl)) + dmns[t] * cA[t].x * ivp;
            const __half nh = __float2half_rn(f);
            *ph = nh;
            *pl = __float2half_rn(f - __half2float(nh));
          } else {
            bf16* ph = reinterpret_cast<bf16*>(SR + off);
            bf16* pl = reinterpret_cast<bf16*>(SR + plane + off);
            const float f = (__bfloat162float(*ph) + __bfloat162float(*pl)) + dmns[t] * cA[t].x * ivp;
            const bf16 nh = __float2bfloat16_rn(f);
            *ph = nh;
            *pl = __float2bfloat16_rn(f - __bfloat162float(nh));
          }
          vq += fixv[t];
        }
      }
      if (prow < NP) vqp[chh * NP + prow] = vq;
    }
    b3_epi_bar();
    for (int i = tid; i < NP; i += 512) vfac[i] = (vqp[i] + vqp[NP + i]) * ivn[i] * ivn[i];
    tc_fence_before();
    fence_proxy_async();
    b3_epi_bar();
    if (lane == 0) mbar_arrive(ds_ready);
    stamp();

    // Everything above needs only the forward's saved buffers and the coefficients; the gradient of the pooled embeddings
    // (dpool_v / dpool_l, written by the kernel launched just before this one) is first read here.  This kernel is launched
    // with the programmatic-dependent-launch attribute, so it may have started while that kernel (and the global InfoNCE
    // backward in front of it) was still running: wait for it now.  (Ordinary stream order otherwise: a no-op.)
    pdl_wait();

    // ---- P4 outputs.  Arithmetic on the accumulators happens with thread = feature column d (TMEM layout), but every
    // global access is TRANSPOSED through a per-warp shared-memory tile [8 rows][32 d]: lane -> (row r = lane / 4, 16-byte
    // chunk ch4 = lane % 4), so one LDG / STG moves 8 rows x 64 contiguous bytes.  (2-byte accesses -- one row per
    // instruction -- are bound by the LSU instruction rate: 64 bytes per warp instruction is the HBM rate per SM.)
    const float invP = 1.f / (float)P;
    const int tw = NT / 4, t_lo = grp * tw;              // multiple of 4
    const int pw = NP / 4, p_lo = grp * pw;              // multiple of 4
    const int dloc = 32 * q + lane;
    const int tr = lane >> 2, tc8 = (lane & 3) * 8;      // transposed role: row inside the chunk, first of 8 columns
    float* stg = reinterpret_cast<float*>(base + L.off_stage) + ew * (8 * 36);
    const int tn = max(0, min(tw, T - t_lo)), pn = max(0, min(pw, P - p_lo));     // live tokens / patches of this group
    const size_t drow = (size_t)32 * q + tc8;            // column offset of this thread's 16-byte piece inside a block
    const bf16* lsrc0 = p.l + ((size_t)b * T + t_lo + tr) * D + drow;
    bf16* ldst0 = p.dl + ((size_t)b * T + t_lo + tr) * D + drow;
    const bf16* vsrc0 = p.v + ((size_t)b * P + p_lo + tr) * D + drow;
    bf16* vdst0 = p.dv + ((size_t)b * P + p_lo + tr) * D + drow;
    // one chunk of 8 (or 4) columns: x[k] (+ per-d term already added) -> tile -> out = x - raw * fac -> global
    // (two halves, so that the NEXT chunk's TMEM load can be issued between them: its ~350-cycle latency then overlaps the
    // transposed read-back, the arithmetic and the store of this chunk)
    auto emit_sts = [&](const float* x) {
#pragma unroll
      for (int k = 0; k < 8; ++k) stg[k * 36 + lane] = x[k];
    };
    auto emit_out = [&](const uint4& raw, float fac, bool row_ok, bf16* dst) {
      __syncwarp();
      const float4 a0 = *reinterpret_cast<const float4*>(stg + tr * 36 + tc8);
      const float4 a1 = *reinterpret_cast<const float4*>(stg + tr * 36 + tc8 + 4);
      __syncwarp();
      float rv[8], ov[8];
      unpack_raw8<kHalf>(raw, rv);
      ov[0] = fmaf(-rv[0], fac, a0.x); ov[1] = fmaf(-rv[1], fac, a0.y); ov[2] = fmaf(-rv[2], fac, a0.z); ov[3] = fmaf(-rv[3], fac, a0.w);
      ov[4] = fmaf(-rv[4], fac, a1.x); ov[5] = fmaf(-rv[5], fac, a1.y); ov[6] = fmaf(-rv[6], fac, a1.z); ov[7] = fmaf(-rv[7], fac, a1.w);
      if (row_ok) *reinterpret_cast<uint4*>(dst) = pack_raw8<kHalf>(ov);
    };
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
    for (int blk = 0; blk < NBLK; ++blk) {
      const size_t dcol = (size_t)blk * 128 + dloc;
      const float dpl = p.dpool_l ? __ldg(p.dpool_l + (size_t)b * D + dcol) : 0.f;
      const float dpv = p.dpool_v ? __ldg(p.dpool_v + (size_t)b * D + dcol) * invP : 0.f;
      {   // unit A: dl[t][d] = dl^T[d][t] (hi-part + lo-part) - l[t][d] lfac_t + m_t dlbar[d] / cnt
        const bf16* lsrc = lsrc0 + blk * 128;
        bf16* ldst = ldst0 + blk * 128;
        uint4 raw[3];
#pragma unroll
        for (int g = 0; g < 3; ++g)
          raw[g] = (8 * g < tw && 8 * g + tr < tn) ? __ldg(reinterpret_cast<const uint4*>(lsrc + (size_t)(8 * g) * D)) : zero4;
        mbar_wait_sleep(oa_full, blk & 1);
        tc_fence_after();
        float xh[8], xl[8];
        if (8 <= tw) { tmem_ld8(tq + cDL + t_lo, xh); tmem_ld8(tq + cDL + NT + t_lo, xl); }
        else { tmem_ld4(tq + cDL + t_lo, xh); tmem_ld4(tq + cDL + NT + t_lo, xl); }
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          const int c = 8 * g;
          if (c < tw) {
            const float4 md0 = *reinterpret_cast<const float4*>(mdl + t_lo + c);
            const float4 md1 = (c + 8 <= tw) ? *reinterpret_cast<const float4*>(mdl + t_lo + c + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            tmem_ld_wait8(xh);
            tmem_ld_wait8(xl);
            const float mdv[8] = {md0.x, md0.y, md0.z, md0.w, md1.x, md1.y, md1.z, md1.w};
            const int nc = (c + 8 <= tw) ? 8 : 4;
            float y[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) y[k] = (k < nc) ? fmaf(mdv[k], dpl, kHalf ? (xh[k] + xl[k]) * isc : xh[k] + xl[k]) : 0.f;
            emit_sts(y);
            if (c + 8 >= tw) {                           // last chunk: the accumulator is in registers
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(oa_free);
            } else if (c + 16 <= tw) {
              tmem_ld8(tq + cDL + t_lo + c + 8, xh); tmem_ld8(tq + cDL + NT + t_lo + c + 8, xl);
            } else {
              tmem_ld4(tq + cDL + t_lo + c + 8, xh); tmem_ld4(tq + cDL + NT + t_lo + c + 8, xl);
            }
            const bool ok = c + tr < tn;
            emit_out(raw[g], ok ? (kHalf ? lfacs[t_lo + c + tr] * isc : lfacs[t_lo + c + tr]) : 0.f, ok, ldst + (size_t)c * D);
          }
        }
        if (blk == 0) stamp();
      }
      {   // unit B: dv[p][d] = dv^T[d][p] - v[p][d] vfac_p + dvbar[d] / P
        const bf16* vsrc = vsrc0 + blk * 128;
        bf16* vdst = vdst0 + blk * 128;
        // raw row pieces are requested two chunks ahead (L2 latency ~ 1 k cycles under load, a chunk takes less); the chunk
        // loop stays ROLLED: one hot body of ~100 instructions instead of 7 copies (instruction fetch)
        uint4 raw0 = (tr < pn) ? __ldg(reinterpret_cast<const uint4*>(vsrc)) : zero4;
        uint4 raw1 = (8 < pw && 8 + tr < pn) ? __ldg(reinterpret_cast<const uint4*>(vsrc + (size_t)8 * D)) : zero4;
        mbar_wait_sleep(ob_full, blk & 1);
        tc_fence_after();
        float x[8];
        if (8 <= pw) tmem_ld8(tq + cDV + p_lo, x);
        else tmem_ld4(tq + cDV + p_lo, x);              // pw is a multiple of 4
        // running pointers / limits instead of per-iteration index arithmetic (the loop overhead was ~20 integer
        // instructions per chunk: ncu source page)
        const bf16* rp = vsrc + (size_t)16 * D;           // raw rows two chunks ahead
        bf16* wp = vdst;
        const float* fp = vfac + p_lo + tr;
        const int rem = pn - tr;                          // this lane's row c0 + tr is live iff c0 < rem
        const int lim2 = min(pw, rem) - 16;               // ... and the row two chunks ahead iff c0 < lim2
        uint32_t ta = tq + cDV + p_lo + 8;
#pragma unroll 1
        for (int c0 = 0; c0 < pw; c0 += 8) {
          const bool full8 = c0 + 8 <= pw;
          const uint4 raw2 = (c0 < lim2) ? __ldg(reinterpret_cast<const uint4*>(rp)) : zero4;
          tmem_ld_wait8(x);
          float y[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) y[k] = (full8 || k < 4) ? (kHalf ? fmaf(x[k], isc, dpv) : x[k] + dpv) : 0.f;
          emit_sts(y);
          if (c0 + 8 >= pw) {                           // last chunk: the accumulator is in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ob_free);
          } else if (c0 + 16 <= pw) {
            tmem_ld8(ta, x);
          } else {
            tmem_ld4(ta, x);
          }
          const bool ok = c0 < rem;
          emit_out(raw0, ok ? (kHalf ? fp[c0] * isc : fp[c0]) : 0.f, ok, wp);
          raw0 = raw1; raw1 = raw2;
          rp += (size_t)8 * D; wp += (size_t)8 * D; ta += 8;
        }
        if (blk == 0) stamp();
      }
    }
    stamp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

bool sparc_bwd3_supported(int P, int T, int D, int dtype) {
  if (dtype != CFA_DTYPE_BF16 && dtype != CFA_DTYPE_F16) return false;
  if (P < 1 || P > 256 || T < 1 || T > 80 || D % 128 || D < 128) return false;
  const Bwd3Layout L = bwd3_layout(P, T, D);
  return L.total + 1024 <= 227 * 1024;
}

int sparc_bwd3_launch(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, float thr, float scale,
                      const float* row_inv_norm, const float* lse_row, const float* lse_col, const float* coef,
                      const float* tt_logits, const float* g_inv_norm, const void* g_split, const float* stats,
                      const float* dpv, const float* dpl, void* dv, void* dl, long long* prof, int dtype, cudaStream_t st,
                      bool pdl_late) {
  if (dtype != CFA_DTYPE_BF16 && dtype != CFA_DTYPE_F16) return CFA_ERR_UNSUPPORTED;
  if (!g_split || !stats || !tt_logits || !g_inv_norm) return CFA_ERR_WORKSPACE;
  const bool half = dtype == CFA_DTYPE_F16;
  const Bwd3Layout L = bwd3_layout(P, T, D);
  CUtensorMap tmV0, tmV1, tmL, tmG;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tmV0, v, D, P, B, 64, L.NP)) != CFA_OK) return rc;
  if ((rc = make_tmap_bf16_3d(&tmV1, v, D, P, B, 64, L.CR0)) != CFA_OK) return rc;     // P4 chunks: NT patch rows
  if ((rc = make_tmap_bf16_3d(&tmL, l, D, T, B, 64, L.NT)) != CFA_OK) return rc;
  if ((rc = make_tmap_bf16_3d(&tmG, g_split, D, T, 2 * (uint64_t)B, 64, L.NT)) != CFA_OK) return rc;
  Bwd3Params prm{prof, P, T, D, thr, scale, mask, row_inv_norm, row_inv_norm + (size_t)B * P, lse_row, lse_col, coef,
                 tt_logits, g_inv_norm, stats, dpv, dpl, g_coef_src ? *g_coef_src : CoefSrc{}, pdl_late ? 1 : 0, (const bf16*)v, (const bf16*)l, (bf16*)dv, (bf16*)dl};
  const size_t smem = L.total + 1024;
#define CFA_B3_LAUNCH(NT_, NP_, D_, H_)                                                                                     \
  do {                                                                                                                      \
    CFA_CUDA_TRY(cudaFuncSetAttribute(sparc_bwd3_kernel<NT_, NP_, D_, H_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    CFA_CUDA_TRY(cfa_launch_pdl(1, sparc_bwd3_kernel<NT_, NP_, D_, H_>, dim3(B), dim3(kB3Threads), smem, st, tmV0, tmV1, tmL, tmG, prm)); \
  } while (0)
  const bool flagship = L.NT == 80 && L.NP == 208 && D == 512;                 // ViT-B/16 (P = 196 / 197, T = 77)
  if (flagship && !half) CFA_B3_LAUNCH(80, 208, 512, false);
  else if (flagship) CFA_B3_LAUNCH(80, 208, 512, true);
  else if (!half) CFA_B3_LAUNCH(0, 0, 0, false);
  else CFA_B3_LAUNCH(0, 0, 0, true);
#undef CFA_B3_LAUNCH
  return launch_status();
}

}  // namespace cfa
