// Restructured tensor-core SPARC backward (tcgen05 / TMEM / TMA) for bf16 embeddings on sm_100a.
//
// Same mathematics as sparc_bwd_tc_kernel (SURVEY.md §8 a-bwd, gradients w.r.t. RAW dot products), reorganised so
// that every contraction is a pure  TMA -> tcgen05.mma  stream and all per-element work happens at three phase
// boundaries only.  The forward saves G (bf16 hi | lo) and Q = G . v^T, which removes every G recomputation and
// every per-block  TMEM -> registers -> shared memory -> MMA  dependency chain from the backward:
//
//   dG = dLhat . l - Gamma G             (Gamma = diag(gfac), the J_n term of normalize(G), losses.py:173)
//   dW = dG . v^T = dLhat . S_raw - Gamma Q                                  (losses.py:245 backward)
//   dv = (dShat + Z)^T . l + (-Gamma W)^T . G - v vfac + dvbar / P           Z = dLhat^T . W
//   dl = (dShat + Z) . v   - l lfac + m dlbar / cnt
//
//   P1   S_raw = l . v^T                       64-wide D blocks, SWIZZLE_128B TMA tiles           (TMEM cS)
//   E1   S -> W (hi/lo), S_raw (hi/lo)         8 epilogue warps: 2 per TMEM lane quarter, each owns half the columns
//   M    dWa = dLhat . S_raw ;  Z = dLhat^T . W                                                    (TMEM cS, cZ)
//   E3   dW = dWa - gfac Q -> renorm / threshold / min-max backward -> dShat' = dShat + Z (hi/lo), Wg = -gfac W (hi/lo)
//   P4a  dl_kb = dShat' . v_kb                 streams v          -> epilogue: - l lfac + pooled term -> global
//   P4b  dv_kb = dShat'^T . l_kb + Wg^T . G_kb streams l, G hi/lo -> epilogue: - v vfac + pooled term -> global
//
// Operands produced on chip stay bf16 hi + lo pairs (~16 mantissa bits), as in the forward.
#include "tc_common.cuh"
#include "sparc_paths.h"
#include <math_constants.h>

namespace cfa {
using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int kB2Threads = 320;                     // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr float kB2MinMaxEps = 1e-8f, kB2ClampEps = 1e-8f;
constexpr uint32_t kB2cS = 0, kB2cZ = 256, kB2cDV = 0, kB2cDL = 256;

struct Bwd2Layout {
  int NP, NT, KB;
  uint32_t l_bytes, v_bytes, slot4, w_bytes, dl_bytes;
  uint32_t off_w, off_s, off_dl, off_f, off_x, off_bar, total;
};

__host__ __device__ inline uint32_t b2_up(uint32_t x, uint32_t a) { return (x + a - 1) / a * a; }

__host__ __device__ inline Bwd2Layout bwd2_layout(int P, int T, int D) {
  Bwd2Layout L;
  L.NP = (P + 15) & ~15; L.NT = (T + 15) & ~15; L.KB = D / 64;
  L.l_bytes = L.NT * 128; L.v_bytes = L.NP * 128;
  L.slot4 = 3 * L.l_bytes > L.v_bytes ? 3 * L.l_bytes : L.v_bytes;      // P4a: one v tile; P4b: l | G hi | G lo tiles
  L.w_bytes = (uint32_t)L.NP * L.NT * 2;                                // one of hi / lo, interleaved [NP/8][NT][8]
  L.dl_bytes = (uint32_t)L.NT * L.NT * 2;
  const uint32_t p1 = L.l_bytes + L.v_bytes;                            // P1 slot: l | v tiles
  const uint32_t ring = b2_up(2 * L.slot4 > p1 ? 2 * L.slot4 : p1, 1024);
  const uint32_t wreg = b2_up(2 * L.w_bytes > p1 ? 2 * L.w_bytes : p1, 1024);        // also P1 slot 1
  const uint32_t sc = (uint32_t)L.NT * (L.NT + 1) * 4;                  // fp32 logits scratch aliases the S_raw region
  const uint32_t sreg = b2_up(2 * L.w_bytes > sc ? 2 * L.w_bytes : sc, 1024);
  L.off_w = ring;
  L.off_s = L.off_w + wreg;
  L.off_dl = L.off_s + sreg;
  const uint32_t p4 = b2_up(2u * (uint32_t)D * 4u, 128) + 8u * 32u * 80u;   // P4: pooled gradients [2][D] + 8 per-warp transpose tiles
  L.off_f = L.off_dl + (2 * L.dl_bytes > p4 ? 2 * L.dl_bytes : p4);
  L.off_x = L.off_f + 4 * (2 * (L.NP + 32) + 7 * L.NT);
  L.off_bar = (L.off_x + 4 * (2 * 2 * L.NT * 4) + 7) & ~7u;              // exchange buffer: [2 sets][2 halves][NT][4]
  L.total = L.off_bar + 8 * 16 + 16;
  // the tensor core reads whole 128-row operand tiles: phantom rows beyond NT / NP must stay inside the allocation
  const uint32_t ntile = L.NP > 128 ? 2 : 1;
  const uint32_t reach = L.off_s + L.w_bytes + ntile * 256 * L.NT + 2048;      // dShat'^T / Wg^T read MN-major
  const uint32_t reach2 = L.off_dl + L.dl_bytes + 256 * L.NT + 2048;            // dLhat^T read MN-major
  if (reach > L.total) L.total = reach;
  if (reach2 > L.total) L.total = reach2;
  return L;
}

struct Bwd2Params {
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
  const float* q_save;      // [B][T][NP]
  const float* dpool_v;
  const float* dpool_l;
  const bf16* v;
  const bf16* l;
  bf16* dv;
  bf16* dl;
};

__device__ __forceinline__ void b2_split8(const float* x, uint4& hi, uint4& lo) { split_hilo8(x, hi, lo); }
__device__ __forceinline__ void b2_unpack8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 b2_pack8(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&t);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void b2_epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Column sums of 16 per-lane values over the 32 lanes of a warp in 31 shuffles (instead of 16 x 5):
// returns, in every lane, the sum over lanes of v[lane & 15].
__device__ __forceinline__ float b2_colsum16(float* v, int lane) {
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

// kNksP / kNksT: NP/16 and NT/16 as compile-time constants (fully unrolled issue loops), or 0 = runtime trip counts.
// kHalf: the raw embeddings (and dv, dl) are fp16 instead of bf16; on-chip operands and the saved G stay bf16 hi/lo.
template <int kNksP, int kNksT, bool kHalf>
__global__ void __launch_bounds__(kB2Threads, 1)
sparc_bwd2_kernel(const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmL,
                  const __grid_constant__ CUtensorMap tmG, const Bwd2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = CFA_SMEM_BASE_1024(smem_raw);
  const Bwd2Layout L = bwd2_layout(p.P, p.T, p.D);
  const int NP = L.NP, NT = L.NT, KB = L.KB, P = p.P, T = p.T, D = p.D;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* ring = base;
  uint8_t* Whi = base + L.off_w;                      // W, later Wg = -gfac W
  uint8_t* Wlo = Whi + L.w_bytes;
  uint8_t* SRhi = base + L.off_s;                     // S_raw, later dShat'
  uint8_t* SRlo = SRhi + L.w_bytes;
  uint8_t* dLhi = base + L.off_dl;
  uint8_t* dLlo = dLhi + L.dl_bytes;
  float* ivn = (float*)(base + L.off_f);              // [NP + 32], zero beyond P
  float* vdot = ivn + NP + 32;                        // [NP + 32]  -> vfac
  float* iln = vdot + NP + 32;                        // [NT]
  float* msk = iln + NT;
  float* lser = msk + NT;
  float* lsec = lser + NT;
  float* ldot = lsec + NT;
  float* lfacs = ldot + NT;
  float* gfacs = lfacs + NT;
  float* xch = (float*)(base + L.off_x);              // [2][2][NT][4]
  uint64_t* bars = (uint64_t*)(base + L.off_bar);
  uint64_t* full = bars;            // [2]
  uint64_t* empty = bars + 2;       // [2]
  uint64_t* s_full = bars + 4;
  uint64_t* e1_ready = bars + 5;
  uint64_t* dl_ready = bars + 6;
  uint64_t* dw_full = bars + 7;
  uint64_t* ds_ready = bars + 8;
  uint64_t* out_full = bars + 9;    // [2]
  uint64_t* out_free = bars + 11;   // [2]
  uint32_t* tmem_slot = (uint32_t*)(bars + 14);

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); mbar_init(out_full + i, 1); mbar_init(out_free + i, 8); }
    mbar_init(s_full, 1); mbar_init(e1_ready, 8); mbar_init(dl_ready, 8); mbar_init(dw_full, 1); mbar_init(ds_ready, 8);
    fence_barrier_init();
    tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmL); tma_prefetch_desc(&tmG);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < 2 * (NP + 32) + 4 * NT; i += kB2Threads) {
    if (i < NP + 32) ivn[i] = (i < P) ? p.inv_vn[(size_t)b * P + i] : 0.f;
    else if (i < 2 * (NP + 32)) vdot[i - NP - 32] = 0.f;
    else {
      const int k = (i - 2 * (NP + 32)) / NT, t = (i - 2 * (NP + 32)) % NT;
      const bool in = t < T;
      float x = 0.f;
      if (k == 0) x = in ? p.inv_ln[(size_t)b * T + t] : 0.f;
      else if (k == 1) x = (in && p.mask[(size_t)b * T + t]) ? 1.f : 0.f;
      else if (k == 2) x = in ? p.lse_row[(size_t)b * T + t] : 0.f;
      else x = in ? p.lse_col[(size_t)b * T + t] : 0.f;
      (k == 0 ? iln : k == 1 ? msk : k == 2 ? lser : lsec)[t] = x;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t il_lbo = (uint32_t)NT * 16;          // interleaved operand: K-major chunk stride == MN-major group stride
  const uint32_t p1_bytes = L.l_bytes + L.v_bytes;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int u = 0;
      for (int kb = 0; kb < KB; ++kb, ++u) {                      // P1: l | v
        const int s = u & 1;
        mbar_wait(empty + s, ((u >> 1) & 1) ^ 1);
        uint8_t* dst = s ? Whi : ring;
        mbar_expect_tx(full + s, p1_bytes);
        tma_load_3d(dst, &tmL, full + s, kb * 64, 0, b);
        tma_load_3d(dst + L.l_bytes, &tmV, full + s, kb * 64, 0, b);
      }
      for (int kb = 0; kb < KB; ++kb, ++u) {                      // P4a: v
        const int s = u & 1;
        mbar_wait(empty + s, ((u >> 1) & 1) ^ 1);
        uint8_t* dst = ring + (size_t)s * L.slot4;
        mbar_expect_tx(full + s, L.v_bytes);
        tma_load_3d(dst, &tmV, full + s, kb * 64, 0, b);
      }
      for (int kb = 0; kb < KB; ++kb, ++u) {                      // P4b: l | G hi | G lo
        const int s = u & 1;
        mbar_wait(empty + s, ((u >> 1) & 1) ^ 1);
        uint8_t* dst = ring + (size_t)s * L.slot4;
        mbar_expect_tx(full + s, 3 * L.l_bytes);
        tma_load_3d(dst, &tmL, full + s, kb * 64, 0, b);
        tma_load_3d(dst + L.l_bytes, &tmG, full + s, kb * 64, 0, 2 * b);
        tma_load_3d(dst + 2 * L.l_bytes, &tmG, full + s, kb * 64, 0, 2 * b + 1);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (warp-uniform control flow, one elected lane issues) ===============================
    const bool leader = elect_one();
    const int nksP = kNksP ? kNksP : NP / 16, nksT = kNksT ? kNksT : NT / 16;
    const uint32_t id_s = make_idesc16(128, NP, false, false, kHalf, kHalf);      // S: raw l (K-major) x raw v (K-major)
    const uint32_t id_dwa = make_idesc16(128, NP, false, true, false, false);     // dLhat (K-major) x S_raw (MN-major), bf16 hi/lo
    const uint32_t id_z = make_idesc16(128, NP, true, true, false, false);        // dLhat^T x W
    const uint32_t id_kn64 = make_idesc16(128, 64, false, true, false, kHalf);    // dShat' (K-major) x raw v tile (MN-major)
    const uint32_t id_nn64 = make_idesc16(128, 64, true, true, false, kHalf);     // dShat'^T (MN-major) x raw l tile (MN-major)
    const uint32_t id_ng64 = make_idesc16(128, 64, true, true, false, false);     // Wg^T (MN-major) x saved G hi/lo tile (bf16)
    const uint64_t sw0 = make_smem_desc(0, 16, 1024, kLayoutSw128);
    auto ilk = [&](const uint8_t* a) { return make_smem_desc(smem_u32(a), il_lbo, 128, kLayoutNone); };   // interleaved, K-major
    auto ilm = [&](const uint8_t* a) { return make_smem_desc(smem_u32(a), 128, il_lbo, kLayoutNone); };   // interleaved, MN-major
    const uint32_t ks_k = (2 * il_lbo) >> 4;          // K-major interleaved: two 8-wide chunks per k-step
    const uint32_t ks_m = 256 >> 4;                   // MN-major interleaved: 16 k-rows of 16 B
    const uint32_t mt_m = il_lbo;                     // MN-major interleaved: second 128-row M tile (16 groups of il_lbo bytes)
    const uint64_t k_dlhi = ilk(dLhi), k_dllo = ilk(dLlo), m_dlhi = ilm(dLhi), m_dllo = ilm(dLlo);
    const uint64_t m_whi = ilm(Whi), m_wlo = ilm(Wlo), m_srhi = ilm(SRhi), m_srlo = ilm(SRlo);
    const uint64_t k_dshi = ilk(SRhi), k_dslo = ilk(SRlo);
    long long* pf = (p.prof && leader) ? p.prof + (size_t)b * 32 : nullptr;
    int pi = 0;
    auto stamp = [&]() { if (pf) pf[pi++] = clock64(); };
    stamp();
    int u = 0;
    // ---- P1: S_raw
    for (int kb = 0; kb < KB; ++kb, ++u) {
      const int s = u & 1;
      mbar_wait(full + s, (u >> 1) & 1);
      tc_fence_after();
      const uint32_t sl = smem_u32(s ? Whi : ring), sv = sl + L.l_bytes;
      const uint64_t dl0 = sw0 | (sl >> 4), dv0 = sw0 | (sv >> 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss_w(leader, tmem + kB2cS, dl0 + 2 * k, dv0 + 2 * k, id_s, (kb | k) != 0);
      umma_commit_w(leader, empty + s);
    }
    umma_commit_w(leader, s_full);
    stamp();
    // ---- dWa = dLhat . S_raw (overwrites S) ; Z = dLhat^T . W
    mbar_wait(dl_ready, 0);
    mbar_wait(e1_ready, 0);
    tc_fence_after();
    stamp();
    // (E1 left -gfac Q in the cS columns: every MMA accumulates, so cS ends up holding dW = dLhat . S_raw - gfac Q)
    _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, tmem + kB2cS, k_dlhi + ks * ks_k, m_srhi + ks * ks_m, id_dwa, true);
    _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, tmem + kB2cS, k_dlhi + ks * ks_k, m_srlo + ks * ks_m, id_dwa, true);
    _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, tmem + kB2cS, k_dllo + ks * ks_k, m_srhi + ks * ks_m, id_dwa, true);
    _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, tmem + kB2cZ, m_dlhi + ks * ks_m, m_whi + ks * ks_m, id_z, ks != 0);
    _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, tmem + kB2cZ, m_dlhi + ks * ks_m, m_wlo + ks * ks_m, id_z, true);
    _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, tmem + kB2cZ, m_dllo + ks * ks_m, m_whi + ks * ks_m, id_z, true);
    umma_commit_w(leader, dw_full);
    // ---- P4a: dl_kb = dShat' . v_kb
    mbar_wait(ds_ready, 0);
    tc_fence_after();
    stamp();
    int o = 0;
    long long wfull = 0, wfree = 0;
    for (int kb = 0; kb < KB; ++kb, ++u, ++o) {
      const int s = u & 1, buf = o & 1;
      const long long w0 = clock64();
      mbar_wait(full + s, (u >> 1) & 1);
      const long long w1 = clock64();
      mbar_wait(out_free + buf, ((o >> 1) & 1) ^ 1);
      wfull += w1 - w0; wfree += clock64() - w1;
      tc_fence_after();
      const uint64_t dv0 = sw0 | (smem_u32(ring + (size_t)s * L.slot4) >> 4);
      const uint32_t d = tmem + kB2cDL + 64 * buf;
      _Pragma("unroll") for (int ks = 0; ks < nksP; ++ks) umma_ss_w(leader, d, k_dshi + ks * ks_k, dv0 + ks * 128, id_kn64, ks != 0);
      _Pragma("unroll") for (int ks = 0; ks < nksP; ++ks) umma_ss_w(leader, d, k_dslo + ks * ks_k, dv0 + ks * 128, id_kn64, true);
      umma_commit_w(leader, out_full + buf);
      umma_commit_w(leader, empty + s);
    }
    stamp();
    if (pf) { pf[8] = wfull; pf[9] = wfree; }
    wfull = 0; wfree = 0;
    // ---- P4b: dv_kb = dShat'^T . l_kb + Wg^T . G_kb
    for (int kb = 0; kb < KB; ++kb, ++u, ++o) {
      const int s = u & 1, buf = o & 1;
      const long long w0 = clock64();
      mbar_wait(full + s, (u >> 1) & 1);
      const long long w1 = clock64();
      mbar_wait(out_free + buf, ((o >> 1) & 1) ^ 1);
      wfull += w1 - w0; wfree += clock64() - w1;
      tc_fence_after();
      const uint32_t sl = smem_u32(ring + (size_t)s * L.slot4);
      const uint64_t dl0 = sw0 | (sl >> 4), gh0 = sw0 | ((sl + L.l_bytes) >> 4), gl0 = sw0 | ((sl + 2 * L.l_bytes) >> 4);
      const int ntile = NP > 128 ? 2 : 1;
      for (int m = 0; m < ntile; ++m) {
        const uint32_t d = tmem + kB2cDV + 128 * buf + 64 * m;
        const uint32_t mo = m * mt_m;
        _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, d, m_srhi + mo + ks * ks_m, dl0 + ks * 128, id_nn64, ks != 0);
        _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, d, m_srlo + mo + ks * ks_m, dl0 + ks * 128, id_nn64, true);
        _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, d, m_whi + mo + ks * ks_m, gh0 + ks * 128, id_ng64, true);
        _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, d, m_whi + mo + ks * ks_m, gl0 + ks * 128, id_ng64, true);
        _Pragma("unroll") for (int ks = 0; ks < nksT; ++ks) umma_ss_w(leader, d, m_wlo + mo + ks * ks_m, gh0 + ks * 128, id_ng64, true);
      }
      umma_commit_w(leader, out_full + buf);
      umma_commit_w(leader, empty + s);
    }
    stamp();
    if (pf) { pf[10] = wfull; pf[11] = wfree; }
  } else {
    // =============================== epilogue: 8 warps, 2 per TMEM lane quarter, each owns half of the columns ===============================
    const int q = warp & 3, h = (warp - 2) >> 2;
    const int row = 32 * q + lane;
    const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
    const bool inT = row < NT;
    const bool valid = row < T && msk[inT ? row : 0] != 0.f;
    const float il = inT ? iln[row] : 0.f;
    const float c_r = p.coef[0], c_c = p.coef[1];
    const int tid = threadIdx.x - 64;                  // 0..255 among the epilogue threads
    float* xme0 = xch + ((0 * 2 + h) * NT + (inT ? row : 0)) * 4;        // exchange slots [set][half][row][4]
    float* xot0 = xch + ((0 * 2 + (1 - h)) * NT + (inT ? row : 0)) * 4;
    const int xset = 2 * NT * 4;
    long long* pf = (p.prof && row == 0 && h == 0) ? p.prof + (size_t)b * 32 + 16 : nullptr;
    int pi = 0;
    auto stamp = [&]() { if (pf) pf[pi++] = clock64(); };
    stamp();

    // ---- phase 0 (overlaps P1): saved T x T logits -> dLhat (hi/lo), gfac_i, ldotL_j
    float* Sc = reinterpret_cast<float*>(SRhi);        // [T][NT+1] fp32 scratch (the S_raw region is free until E1)
    const int ldl = NT + 1;
    {
      // 24 KB per sample, L2-cold at this point (the forward wrote it a whole kernel ago): issue the loads 12 deep per
      // thread instead of one dependent round trip per element — phase 0 gates E1 (same warps) and was 21-25 k cycles
      const float* src = p.tt_logits + (size_t)b * T * T;
      const int n = T * T;
      for (int base = 0; base < n; base += 256 * 12) {
        float tmp[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) { const int idx = base + tid + 256 * k; tmp[k] = idx < n ? __ldg(src + idx) : 0.f; }
#pragma unroll
        for (int k = 0; k < 12; ++k) {
          const int idx = base + tid + 256 * k;
          if (idx < n) { const int i = idx / T, j = idx - i * T; Sc[i * ldl + j] = tmp[k]; }
        }
      }
    }
    const float ign = (row < T) ? p.g_inv_norm[(size_t)b * T + row] : 0.f;
    b2_epi_bar();
    {
      const int tsplit = ((NT / 16 + 1) / 2) * 16;
      const int t_lo = h ? tsplit : 0, t_hi = h ? NT : tsplit;
      const float lr = inT ? lser[row] : 0.f;
      float gdot = 0.f;
      for (int c0 = t_lo; c0 < t_hi; c0 += 16) {
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
        if (inT) {
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            uint4 hi, lo;
            b2_split8(x + 8 * g8, hi, lo);
            const uint32_t off = il_offset(NT, row, c0 + 8 * g8);
            *reinterpret_cast<uint4*>(dLhi + off) = hi;
            *reinterpret_cast<uint4*>(dLlo + off) = lo;
          }
        }
      }
      if (inT) xme0[0] = gdot;
      b2_epi_bar();                                    // exchange set 0; also: every pr is in Sc
      if (inT && h == 0) gfacs[row] = (gdot + xot0[0]) * ign * ign;      // (g^_i . dg^_i) / ||G_i||^2
      if (h == 1 && inT) {                             // the other half of the warps forms the column sums
        float s = 0.f;
        if (row < T) for (int i = 0; i < T; ++i) s += Sc[i * ldl + row];
        ldot[row] = s;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(dl_ready);
    }
    stamp();

    // ---- E1: S -> W (hi/lo), S_raw (hi/lo), row statistics
    const int psplit = ((NP / 16 + 1) / 2) * 16;
    const int c_lo = h ? psplit : 0, c_hi = h ? NP : psplit;
    mbar_wait(s_full, 0);
    tc_fence_after();
    stamp();
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    int imn = 0x7fffffff, imx = 0x7fffffff;
    for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
      float x[16];
      tmem_ld16(trow + kB2cS + c0, x);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float s = x[j] * il * ivn[c0 + j];
        const bool in = c0 + j < P;
        const bool lt = in && s < mn, gt = in && s > mx;      // strict: first occurrence wins, like torch.min/max
        mn = lt ? s : mn; imn = lt ? c0 + j : imn;
        mx = gt ? s : mx; imx = gt ? c0 + j : imx;
      }
    }
    {
      float* xme = xme0 + xset;                        // exchange set 1
      float* xot = xot0 + xset;
      if (inT) { xme[0] = mn; xme[1] = __int_as_float(imn); xme[2] = mx; xme[3] = __int_as_float(imx); }
      b2_epi_bar();
      if (inT) {
        const float omn = xot[0], omx = xot[2];
        const int oimn = __float_as_int(xot[1]), oimx = __float_as_int(xot[3]);
        if (omn < mn || (omn == mn && oimn < imn)) { mn = omn; imn = oimn; }
        if (omx > mx || (omx == mx && oimx < imx)) { mx = omx; imx = oimx; }
      }
    }
    const float rng = mx - mn + kB2MinMaxEps;
    const float inv_rng = 1.f / rng;
    float sum = 0.f;
    for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
      float x[16];
      tmem_ld16(trow + kB2cS + c0, x);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float nn = (x[j] * il * ivn[c0 + j] - mn) * inv_rng;
        sum += (c0 + j < P && !(nn < p.thr)) ? nn : 0.f;
      }
    }
    if (inT) xme0[0] = sum;                            // exchange set 0 again (its previous readers passed the set-1 barrier)
    b2_epi_bar();
    if (inT) sum += xot0[0];
    const float sigma = fmaxf(sum, kB2ClampEps);
    const float inv_sigma = valid ? 1.f / sigma : 0.f;
    // sweep 3 also parks -gfac Q in the S columns it has just consumed: the dWa MMAs then accumulate on top of it,
    // so the backward never touches Q again (its L2 latency hides behind the two hi/lo splits of the previous chunk)
    const float gfac = inT ? gfacs[row] : 0.f;          // visible: written before the barriers above
    const float* qrow = p.q_save + ((size_t)b * T + (row < T ? row : 0)) * NP;
    auto load_q16 = [&](int c, float* qv) {
#pragma unroll
      for (int g4 = 0; g4 < 4; ++g4) {
        const float4 t4 = __ldg(reinterpret_cast<const float4*>(qrow + c) + g4);
        qv[4 * g4] = t4.x; qv[4 * g4 + 1] = t4.y; qv[4 * g4 + 2] = t4.z; qv[4 * g4 + 3] = t4.w;
      }
    };
    {
      float qc[16], qn[16];
      if (c_lo < c_hi) load_q16(c_lo, qc);
      for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
        float x[16], w[16];
        tmem_ld16(trow + kB2cS + c0, x);
        if (c0 + 16 < c_hi) load_q16(c0 + 16, qn);
        tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 iv = *reinterpret_cast<const float4*>(ivn + c0 + 4 * j4);
          const float ivv[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = 4 * j4 + jj;
            const float nn = (x[j] * il * ivv[jj] - mn) * inv_rng;
            const bool in = valid && c0 + j < P;
            w[j] = (in && !(nn < p.thr)) ? nn * inv_sigma : 0.f;
            x[j] = in ? x[j] : 0.f;
            qc[j] = valid ? -gfac * qc[j] : 0.f;
          }
        }
        if (inT) {
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            const uint32_t off = il_offset(NT, row, c0 + 8 * g8);
            uint4 hi, lo;
            b2_split8(w + 8 * g8, hi, lo);
            *reinterpret_cast<uint4*>(Whi + off) = hi;
            *reinterpret_cast<uint4*>(Wlo + off) = lo;
            b2_split8(x + 8 * g8, hi, lo);
            *reinterpret_cast<uint4*>(SRhi + off) = hi;
            *reinterpret_cast<uint4*>(SRlo + off) = lo;
          }
        }
        tmem_st16(trow + kB2cS + c0, qc);
#pragma unroll
        for (int j = 0; j < 16; ++j) qc[j] = qn[j];
      }
      tmem_st_wait();
    }
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(e1_ready);
    stamp();

    // ---- E3: dW = dWa - gfac Q -> renorm / threshold / min-max backward -> dShat' (hi/lo), Wg (hi/lo), lfac, vfac
    const float isg = 1.f / sigma;
    const bool keep_all = p.thr <= 0.f;
    auto load_w16 = [&](int c, float* w) {               // this row's W[c .. c+16) = hi + lo
#pragma unroll
      for (int g8 = 0; g8 < 2; ++g8) {
        float hh[8], ll[8];
        const uint32_t off = il_offset(NT, inT ? row : 0, c + 8 * g8);
        b2_unpack8(*reinterpret_cast<const uint4*>(Whi + off), hh);
        b2_unpack8(*reinterpret_cast<const uint4*>(Wlo + off), ll);
#pragma unroll
        for (int j = 0; j < 8; ++j) w[8 * g8 + j] = hh[j] + ll[j];
      }
    };
    mbar_wait(dw_full, 0);
    tc_fence_after();
    stamp();
    // Renorm / threshold / min-max backward in ONE sweep.  Two identities make that possible:
    //   sum_p W dW = dG . G = 0   (dG = J_n(G)^T dg^ is orthogonal to G), so  dTheta = dW / sigma  without a prior row sum;
    //   sum_p dN N = 0            (W is invariant to the scale of N), so only the arg-MIN element receives a scatter
    //                             term, dmn = -(sum_kept dW) / (sigma rng), which is patched in after the sweep.
    float* vq = reinterpret_cast<float*>(dLhi);         // [4 quarters][NP] column partials (dLhat is dead: dWa / Z are done)
    float cx = 0.f, sdot = 0.f;
    for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
      float x[16], w[16], z[16], pr[16];
      tmem_ld16(trow + kB2cS + c0, x);
      tmem_ld16(trow + kB2cZ + c0, z);
      load_w16(c0, w);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int pc = c0 + j;
        const bool in = valid && pc < P;
        const bool kept = (keep_all || w[j] > 0.f) && in;
        const float dw = kept ? x[j] : 0.f;              // phantom rows may hold NaN/Inf: select, never 0*x
        cx += dw;
        const float ds = (dw * isg) * inv_rng;
        const float sv = fmaf(w[j] * sigma, rng, mn);    // normalised similarity of a kept element
        const float prod = kept ? ds * sv : 0.f;         // select, never 0 * x (phantom rows hold garbage statistics)
        sdot += prod;
        pr[j] = prod;
        x[j] = in ? fmaf(ds * il, ivn[pc], z[j]) : 0.f;    // dShat' = dShat + Z
        w[j] = in ? -gfac * w[j] : 0.f;                    // Wg
      }
      if (inT) {
#pragma unroll
        for (int g8 = 0; g8 < 2; ++g8) {
          const uint32_t off = il_offset(NT, row, c0 + 8 * g8);
          uint4 hi, lo;
          b2_split8(x + 8 * g8, hi, lo);
          *reinterpret_cast<uint4*>(SRhi + off) = hi;
          *reinterpret_cast<uint4*>(SRlo + off) = lo;
          b2_split8(w + 8 * g8, hi, lo);
          *reinterpret_cast<uint4*>(Whi + off) = hi;
          *reinterpret_cast<uint4*>(Wlo + off) = lo;
        }
      }
      const float cs = b2_colsum16(pr, lane);           // column sums over this warp's 32 rows
      if (lane < 16) vq[q * NP + c0 + lane] = cs;       // per-quarter partial (fixed-order sum below: deterministic bits)
    }
    stamp();
    int* fixi = reinterpret_cast<int*>(lser);           // [NT] arg-min column of each row (-1: none); lser / lsec are dead
    float* fixv = lsec;                                 // [NT] its contribution to the column sum
    {
      float* xme = xme0 + xset;                        // exchange set 1
      float* xot = xot0 + xset;
      if (inT) { xme[0] = cx; xme[1] = sdot; }
      b2_epi_bar();
      if (inT) { cx += xot[0]; sdot += xot[1]; }
    }
    const float dmn = valid ? -cx * isg * inv_rng : 0.f;
    if (valid && imn >= c_lo && imn < c_hi) {           // the half that owns the arg-min column patches dShat'[t][imn]
      const uint32_t off = il_offset(NT, row, imn);
      bf16* ph = reinterpret_cast<bf16*>(SRhi + off);
      bf16* pl = reinterpret_cast<bf16*>(SRlo + off);
      const float f = (__bfloat162float(*ph) + __bfloat162float(*pl)) + dmn * il * ivn[imn];
      const bf16 nh = __float2bfloat16_rn(f);
      *ph = nh;
      *pl = __float2bfloat16_rn(f - __bfloat162float(nh));
    }
    if (inT && h == 0) {
      fixi[row] = valid ? imn : -1;
      fixv[row] = dmn * mn;                             // ds[imn] * s[imn], s[imn] = mn
      lfacs[row] = (sdot + dmn * mn + ldot[row]) * il * il;        // (l^_t . dl^_t) / ||l_t||^2
    }
    stamp();
    b2_epi_bar();
    stamp();
    for (int i = tid; i < NP; i += 256) {               // -> vfac_p
      float acc = (vq[i] + vq[NP + i]) + (vq[2 * NP + i] + vq[3 * NP + i]);
      for (int t = 0; t < T; ++t) acc += (fixi[t] == i) ? fixv[t] : 0.f;
      vdot[i] = acc * ivn[i] * ivn[i];
    }
    b2_epi_bar();
    tc_fence_before();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(ds_ready);
    stamp();

    // ---- P4 outputs: each warp owns 32 of the 64 columns of a block
    float cnt = 0.f;
    for (int t = 0; t < T; ++t) cnt += msk[t];
    const float invc = 1.f / fmaxf(cnt, kB2ClampEps), invP = 1.f / (float)P;
    const float mrow = inT ? msk[row] * invc : 0.f;
    const float lfac = inT ? lfacs[row] : 0.f;
    const float vf0 = (row < P) ? vdot[row] : 0.f, vf1 = (128 + row < P) ? vdot[128 + row] : 0.f;
    int o = 0;
    // pooled-mean gradients of this sample -> shared memory (the dLhat region is dead once dWa / Z are done)
    float* dps = reinterpret_cast<float*>(dLhi);         // [2][D]: dvbar, dlbar
    for (int i = tid; i < 2 * D; i += 256) {
      const float* src = (i < D) ? p.dpool_v : p.dpool_l;
      dps[i] = src ? __ldg(src + (size_t)b * D + (i < D ? i : i - D)) : 0.f;
    }
    b2_epi_bar();
    // Raw rows (needed for the - x fac terms) are fetched one block AHEAD of the accumulator they are combined with, and
    // every global access of the epilogue is TRANSPOSED through a per-warp shared-memory tile: thread = row for the
    // arithmetic (TMEM layout), but lane -> (row = 8 it + lane / 4, 16-byte chunk = lane % 4) for LDG / STG, so one
    // instruction touches 8 rows x 64 contiguous bytes instead of 32 rows x 16 bytes (4x fewer L1 wavefronts: the
    // output phases were LSU-bound, not tensor-bound).
    uint8_t* sc = dLhi + b2_up(2u * (uint32_t)D * 4u, 128) + (warp - 2) * (32 * 80);
    const int tr = lane >> 2, tc16 = (lane & 3) * 16, tc8 = (lane & 3) * 8;
    auto load_raw = [&](const bf16* gtile, int nrows, uint4* r4) {
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int r = it * 8 + tr;
        r4[it] = (r < nrows) ? __ldg(reinterpret_cast<const uint4*>(gtile + (size_t)r * D + tc8)) : make_uint4(0, 0, 0, 0);
      }
    };
    auto emit = [&](const float* x, const uint4* rawT, const float* dp, float fac, float dscale, bf16* gtile, int nrows) {
#pragma unroll
      for (int it = 0; it < 4; ++it) *reinterpret_cast<uint4*>(sc + (it * 8 + tr) * 80 + tc16) = rawT[it];
      __syncwarp();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float rv[8], ov[8];
        unpack_raw8<kHalf>(*reinterpret_cast<const uint4*>(sc + lane * 80 + g * 16), rv);
        const float4 d0v = *reinterpret_cast<const float4*>(dp + 8 * g), d1v = *reinterpret_cast<const float4*>(dp + 8 * g + 4);
        const float dpv[8] = {d0v.x, d0v.y, d0v.z, d0v.w, d1v.x, d1v.y, d1v.z, d1v.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) ov[j] = fmaf(dpv[j], dscale, fmaf(-rv[j], fac, x[8 * g + j]));
        *reinterpret_cast<uint4*>(sc + lane * 80 + g * 16) = pack_raw8<kHalf>(ov);
      }
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int r = it * 8 + tr;
        const uint4 ov4 = *reinterpret_cast<const uint4*>(sc + r * 80 + tc16);
        if (r < nrows) *reinterpret_cast<uint4*>(gtile + (size_t)r * D + tc8) = ov4;
      }
      __syncwarp();
    };
    {
      const int nrows = min(32, max(0, T - 32 * q));
      const bf16* lt = p.l + ((size_t)b * T + 32 * q) * D + 32 * h;
      bf16* dlt = p.dl + ((size_t)b * T + 32 * q) * D + 32 * h;
      uint4 raw[4];
      load_raw(lt, nrows, raw);
      for (int kb = 0; kb < KB; ++kb, ++o) {               // dl
        const int buf = o & 1, d0 = kb * 64 + 32 * h;
        uint4 rawn[4];
        if (kb + 1 < KB) load_raw(lt + (kb + 1) * 64, nrows, rawn);
        mbar_wait(out_full + buf, (o >> 1) & 1);
        tc_fence_after();
        float x[32];
        tmem_ld32(trow + kB2cDL + 64 * buf + 32 * h, x);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(out_free + buf);        // the TMEM buffer is free as soon as it is in registers
        if (nrows > 0) emit(x, raw, dps + D + d0, lfac, mrow, dlt + kb * 64, nrows);
#pragma unroll
        for (int g = 0; g < 4; ++g) raw[g] = rawn[g];
      }
    }
    stamp();
    {
      const int ntile = NP > 128 ? 2 : 1;
      const int nr0 = min(32, max(0, P - 32 * q)), nr1 = ntile > 1 ? min(32, max(0, P - 128 - 32 * q)) : 0;
      const bf16* vt0 = p.v + ((size_t)b * P + 32 * q) * D + 32 * h;
      const bf16* vt1 = vt0 + (size_t)128 * D;
      bf16* dvt0 = p.dv + ((size_t)b * P + 32 * q) * D + 32 * h;
      bf16* dvt1 = dvt0 + (size_t)128 * D;
      uint4 raw0[4], raw1[4];
      load_raw(vt0, nr0, raw0);
      load_raw(vt1, nr1, raw1);
      for (int kb = 0; kb < KB; ++kb, ++o) {               // dv
        const int buf = o & 1, d0 = kb * 64 + 32 * h;
        uint4 raw0n[4], raw1n[4];
        if (kb + 1 < KB) { load_raw(vt0 + (kb + 1) * 64, nr0, raw0n); load_raw(vt1 + (kb + 1) * 64, nr1, raw1n); }
        mbar_wait(out_full + buf, (o >> 1) & 1);
        tc_fence_after();
        float x[32];
        tmem_ld32(trow + kB2cDV + 128 * buf + 32 * h, x);
        tmem_ld_wait();
        if (nr0 > 0) emit(x, raw0, dps + d0, vf0, invP, dvt0 + kb * 64, nr0);
        tmem_ld32(trow + kB2cDV + 128 * buf + 64 + 32 * h, x);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(out_free + buf);
        if (nr1 > 0) emit(x, raw1, dps + d0, vf1, invP, dvt1 + kb * 64, nr1);
#pragma unroll
        for (int g = 0; g < 4; ++g) { raw0[g] = raw0n[g]; raw1[g] = raw1n[g]; }
      }
    }
    stamp();
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

bool sparc_bwd2_supported(int P, int T, int D, int dtype) {
  if (dtype != CFA_DTYPE_BF16) return false;           // see sparc_fwd2_supported: mixed fp16 x bf16 MMAs are illegal
  if (!sparc_tc_supported(P, T, D, CFA_DTYPE_BF16)) return false;
  const Bwd2Layout L = bwd2_layout(P, T, D);
  if (L.NP > 256 || L.NT > 128 || D % 64) return false;
  return L.total + 1024 <= 227 * 1024;
}

int sparc_bwd2_launch(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, float thr, float scale,
                      const float* row_inv_norm, const float* lse_row, const float* lse_col, const float* coef,
                      const float* tt_logits, const float* g_inv_norm, const void* g_split, const float* q_save,
                      const float* dpv, const float* dpl, void* dv, void* dl, long long* prof, int dtype, cudaStream_t st) {
  const bool half = dtype == CFA_DTYPE_F16;
  if (half) return CFA_ERR_UNSUPPORTED;
  const Bwd2Layout L = bwd2_layout(P, T, D);
  CUtensorMap tmV, tmL, tmG;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tmV, v, D, P, B, 64, L.NP, half)) != CFA_OK) return rc;
  if ((rc = make_tmap_bf16_3d(&tmL, l, D, T, B, 64, L.NT, half)) != CFA_OK) return rc;
  if ((rc = make_tmap_bf16_3d(&tmG, g_split, D, T, 2 * (uint64_t)B, 64, L.NT)) != CFA_OK) return rc;
  Bwd2Params prm{prof, P, T, D, thr, scale, mask, row_inv_norm, row_inv_norm + (size_t)B * P, lse_row, lse_col, coef,
                 tt_logits, g_inv_norm, q_save, dpv, dpl, (const bf16*)v, (const bf16*)l, (bf16*)dv, (bf16*)dl};
  const size_t smem = L.total + 1024;
#define CFA_B2_LAUNCH(NP_, NT_, HALF)                                                                                     \
  do {                                                                                                                    \
    CFA_CUDA_TRY(cudaFuncSetAttribute(sparc_bwd2_kernel<NP_, NT_, HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    sparc_bwd2_kernel<NP_, NT_, HALF><<<B, kB2Threads, smem, st>>>(tmV, tmL, tmG, prm);                                   \
  } while (0)
  if (L.NP == 208 && L.NT == 80) {          // ViT-B/16 (P = 196 / 197, T = 77): fully unrolled issue loops
    CFA_B2_LAUNCH(13, 5, false);
  } else {
    CFA_B2_LAUNCH(0, 0, false);
  }
#undef CFA_B2_LAUNCH
  return launch_status();
}

}  // namespace cfa
