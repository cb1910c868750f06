// Tensor-core global InfoNCE (tcgen05 / TMEM / TMA): local rows x (all-gathered) global columns, both directions.
//
//   split kernel : x_hat = x / max(|x|, eps) for every row, stored as bf16 hi + lo (x_hat ~ hi + lo, ~16 mantissa bits)
//   forward      : one CTA per 128 x 128 logits tile; S = A_hi B_hi^T + A_hi B_lo^T + A_lo B_hi^T accumulated in TMEM
//                  over 64-wide D blocks (TMA, SWIZZLE_128B); epilogue = streaming max / sum-exp per row (+ target logit);
//                  the B x Bg logits never reach memory (losses.py:160-162, :21-28)
//   backward     : recompute the tile, dS = c_self softmax_row + c_other softmax_col - (c_self + c_other) I as bf16 hi/lo
//                  in shared memory (A operand), then dA_hat[128 x D] += dS . B_hat (B tiles read MN-major) in TMEM
// Warp roles as in sparc_tc_fwd2.cu: warp 0 TMA, warp 1 MMA issue (warp-uniform, elected lane), warps 2-9 epilogue
// (two per TMEM lane quarter, each owning half of the tile columns: the epilogues are latency-bound, not throughput-bound).
#include "tc_common.cuh"
#include <math_constants.h>

namespace cfa {
using namespace tc;
typedef __nv_bfloat16 bf16;

constexpr int kGtThreads = 320;          // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue: 2 per TMEM lane quarter, half the columns each
constexpr int kGtM = 128, kGtN = 128;
constexpr uint32_t kGtTileA = kGtM * 128, kGtTileB = kGtN * 128;                 // one 64-wide bf16 block of a tile
constexpr uint32_t kGtStage = 2 * kGtTileA + 2 * kGtTileB;                       // A_hi A_lo B_hi B_lo = 64 KB
constexpr int kGtStages = 3;          // backward (one CTA per SM: it holds all 512 TMEM columns) and small forward grids
constexpr int kGtFwdStages = 1;       // large forward grids: ONE 64 KB stage per CTA and three CTAs per SM -- the CTAs interleave: three loads in flight per SM
                                      // and one tile's log-sum-exp epilogue under the other tiles' MMAs (2 stages, 1 CTA/SM: 123 us per launch at
                                      // 1024 x 8192; this way: see DESIGN.md 3.4)

// ---------------------------------------------------------------------------------------------------------------
// normalise + split: out[which][row][d], which = 0: a_hi, 1: a_lo, 2: b_hi, 3: b_lo
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gt_split_kernel(const float* __restrict__ a, const float* __restrict__ b, int rows, int D, float eps,
                bf16* __restrict__ out, float* __restrict__ norms /* [2][loc_rows] or NULL */, int rank_rows, size_t rank_stride,
                const PeerTable peers, int loc_row0, int loc_rows) {
  pdl_launch_dependents();                             // programmatic dependent launch (common.cuh): start early, wait here
  pdl_wait();
  // norms are kept for the LOCAL rows only (global rows [loc_row0, loc_row0 + loc_rows)): the local operand tiles are a
  // row window of the all-rows array, so one launch serves both
  // rank_rows / rank_stride: global row g = r * rank_rows + i lives at base + r * rank_stride + i * D (the raw output of
  // an all-gather of per-rank [2][B][D] blocks); rank_rows == rows, rank_stride == 0 for a plain [rows][D] matrix
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, which = blockIdx.y;
  if (r >= rows) return;
  // peers.n > 0: rank q's [2][rank_rows][D] block is read where it lives, in q's HBM over NVLink (the gather is fused
  // into this, its only consumer)
  const float* x = peers.n ? peers.base[r / rank_rows] + ((size_t)which * rank_rows + (size_t)(r % rank_rows)) * D
                           : (which ? b : a) + (size_t)(r / rank_rows) * rank_stride + (size_t)(r % rank_rows) * D;
  // the row is read ONCE (it may live in a peer's HBM): D <= 1024 -> up to 32 values per lane stay in registers
  float xr[32];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const int d = lane + 32 * k;
    xr[k] = d < D ? x[d] : 0.f;
    ss = fmaf(xr[k], xr[k], ss);
  }
  ss = warp_sum(ss);
  const float n = fmaxf(sqrtf(ss), eps);
  if (norms && lane == 0 && r >= loc_row0 && r < loc_row0 + loc_rows) norms[(size_t)which * loc_rows + (r - loc_row0)] = n;
  bf16* hi = out + ((size_t)(2 * which) * rows + r) * D;
  bf16* lo = out + ((size_t)(2 * which + 1) * rows + r) * D;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const int d = lane + 32 * k;
    if (d < D) {
      const float y = xr[k] / n;
      const bf16 h = __float2bfloat16_rn(y);
      hi[d] = h;
      lo[d] = __float2bfloat16_rn(y - __bfloat162float(h));
    }
  }
}

struct GtParams {
  int B, Bg, D, col_offset;
  float scale;
  // forward
  float* part_m; float* part_l; float* diag;
  // backward
  const float* lse_loc; const float* lse_all; const float* coef; float* dpart;
  CoefSrc cs;                           // cs.on: coefficients from the upstream-gradient pointers (one-call entry points)
  int dsplit;                           // backward: the D / 64 output blocks are divided over this many CTAs (grid z = 2 dsplit)
  int dpart_atomic;                     // != 0: dpart is ONE zeroed [2][B][D] accumulator, column tiles add into it (red.global)
  int lse_rank_rows, lse_rank_stride;   // > 0: lse_all is the raw all-gather of per-rank [lse_a | lse_b | 2 sums] packs
  volatile int* dbg;      // optional host-mapped progress markers (cfa_debug_set_marker_buffer), [cta][16 warps]
};
static int* g_gt_dbg = nullptr;
thread_local const CoefSrc* g_coef_src = nullptr;      // set by the one-call entry points around their backward (sparc_paths.h)
#define GT_MARK(v) do { if (p.dbg && lane == 0) { p.dbg[((((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + warp) * 32) + (v)] = (int)(clock64() - gt_t0); } } while (0)


struct GtSmem {
  uint8_t* stages;       // kGtStages x kGtStage (forward / recompute); aliased by dS hi/lo + output-phase stages in bwd
  uint64_t* bars;
  uint32_t* tmem_slot;
};

// S tile (128 x 128) into TMEM columns [0,128): TMA producer + MMA issuer roles; epilogue warps just wait on s_full.
template <int kStages>
__device__ __forceinline__ void gt_issue_logits(int warp, int lane, uint8_t* stages, uint64_t* full, uint64_t* empty,
                                                uint64_t* s_full, uint32_t tmem, const CUtensorMap* tmR, const CUtensorMap* tmC,
                                                int row0, int col0, int ra, int ca, int KB) {
  if (warp == 0) {
    if (lane == 0) {
      for (int u = 0; u < KB; ++u) {
        const int slot = u % kStages;
        mbar_wait(empty + slot, ((u / kStages) & 1) ^ 1);
        uint8_t* st = stages + (size_t)slot * kGtStage;
        mbar_expect_tx(full + slot, kGtStage);
        tma_load_3d(st, tmR, full + slot, u * 64, row0, ra);                          // A_hi
        tma_load_3d(st + kGtTileA, tmR, full + slot, u * 64, row0, ra + 1);           // A_lo
        tma_load_3d(st + 2 * kGtTileA, tmC, full + slot, u * 64, col0, ca);           // B_hi
        tma_load_3d(st + 2 * kGtTileA + kGtTileB, tmC, full + slot, u * 64, col0, ca + 1);   // B_lo
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(kGtM, kGtN, false, false);
    const uint64_t sw0 = make_smem_desc(0, 16, 1024, kLayoutSw128);
    for (int u = 0; u < KB; ++u) {
      const int slot = u % kStages;
      mbar_wait(full + slot, (u / kStages) & 1);
      tc_fence_after();
      const uint32_t s0 = smem_u32(stages + (size_t)slot * kGtStage);
      const uint64_t ahi = sw0 | (s0 >> 4), alo = sw0 | ((s0 + kGtTileA) >> 4);
      const uint64_t bhi = sw0 | ((s0 + 2 * kGtTileA) >> 4), blo = sw0 | ((s0 + 2 * kGtTileA + kGtTileB) >> 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        umma_ss_w(leader, tmem, ahi + 2 * k, bhi + 2 * k, idesc, (u | k) != 0);
        umma_ss_w(leader, tmem, ahi + 2 * k, blo + 2 * k, idesc, true);
        umma_ss_w(leader, tmem, alo + 2 * k, bhi + 2 * k, idesc, true);
      }
      umma_commit_w(leader, empty + slot);
    }
    umma_commit_w(leader, s_full);
  }
}

template <int kStages>
__global__ void __launch_bounds__(kGtThreads, kStages == 1 ? 3 : 1)
gt_fwd_kernel(const __grid_constant__ CUtensorMap tmLoc, const __grid_constant__ CUtensorMap tmAll, const GtParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stages = base;
  uint64_t* bars = (uint64_t*)(base + kStages * kGtStage);
  uint64_t* full = bars; uint64_t* empty = bars + kStages; uint64_t* s_full = bars + 2 * kStages;
  uint32_t* tmem_slot = (uint32_t*)(s_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dir = blockIdx.z, row0 = blockIdx.x * kGtM, ct = blockIdx.y, col0 = ct * kGtN, nct = gridDim.y;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(s_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmLoc); tma_prefetch_desc(&tmAll);
  }
  pdl_launch_dependents();
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();                                          // prologue done under the split kernel; its output is read from here on
  gt_issue_logits<kStages>(warp, lane, stages, full, empty, s_full, tmem, &tmLoc, &tmAll, row0, col0, dir ? 2 : 0, dir ? 0 : 2, p.D / 64);
  if (warp >= 2) {
    const int q = warp & 3, h = (warp - 2) >> 2, row = 32 * q + lane, grow = row0 + row;
    const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
    mbar_wait(s_full, 0);
    tc_fence_after();
    float m = -CUDART_INF_F, l = 0.f;
    for (int c0 = 64 * h; c0 < 64 * h + 64; c0 += 32) {
      float x[32];
      tmem_ld32(trow + c0, x);
      tmem_ld_wait();
      float cm = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int gcol = col0 + c0 + j;
        x[j] = (gcol < p.Bg) ? x[j] * p.scale : -CUDART_INF_F;
        cm = fmaxf(cm, x[j]);
        if (grow < p.B && gcol == p.col_offset + grow) p.diag[(size_t)dir * p.B + grow] = x[j];
      }
      const float nm = fmaxf(m, cm);
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) s += (x[j] == -CUDART_INF_F) ? 0.f : __expf(x[j] - nm);
      l = l * ((m == -CUDART_INF_F) ? 0.f : __expf(m - nm)) + s;
      m = nm;
    }
    if (grow < p.B) {                                  // one online-softmax partial per (column tile, half)
      p.part_m[((size_t)dir * 2 * nct + 2 * ct + h) * p.B + grow] = m;
      p.part_l[((size_t)dir * 2 * nct + 2 * ct + h) * p.B + grow] = l;
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

// ---------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gt_split8(const float* x, uint4& hi, uint4& lo) {
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

constexpr uint32_t kGtDs = kGtM * kGtN * 2;          // dS hi (or lo) tile, interleaved [k/8][row][8]: 32 KB
constexpr uint32_t kGtOutStage = 2 * kGtTileB;       // B_hi, B_lo 64-wide block for the output contraction: 32 KB

__global__ void __launch_bounds__(kGtThreads, 1)
gt_bwd_kernel(const __grid_constant__ CUtensorMap tmLoc, const __grid_constant__ CUtensorMap tmAll, const GtParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stages = base;                              // phase A: 3 x 64 KB
  uint8_t* dShi = base;                                // phase B/C alias: dS hi | dS lo | 2 output stages
  uint8_t* dSlo = base + kGtDs;
  uint8_t* ostg = base + 2 * kGtDs;
  uint64_t* bars = (uint64_t*)(base + kGtStages * kGtStage);
  uint64_t* full = bars; uint64_t* empty = bars + 3; uint64_t* s_full = bars + 6; uint64_t* ds_ready = bars + 7;
  uint64_t* ofull = bars + 8; uint64_t* oempty = bars + 10; uint64_t* o_done = bars + 12;
  uint32_t* tmem_slot = (uint32_t*)(bars + 13);
  static_assert(kGtStages == 3, "barrier layout above");
  float* lse_s = (float*)(bars + 16);                  // [kGtN] the other direction's LSE of this tile's columns
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dir = blockIdx.z & 1, dz = blockIdx.z >> 1, row0 = blockIdx.x * kGtM, ct = blockIdx.y, col0 = ct * kGtN, nct = gridDim.y;
  const int D = p.D, KB = D / 64;
  // output blocks of this CTA (small problems: the D / 64 blocks are divided over dsplit CTAs, each recomputes the logits)
  const int u_lo = dz * KB / p.dsplit, u_hi = (dz + 1) * KB / p.dsplit;
  const long long gt_t0 = clock64();
  const int ra = dir ? 2 : 0, ca = dir ? 0 : 2;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kGtStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(ofull + i, 1); mbar_init(oempty + i, 1); }
    mbar_init(s_full, 1); mbar_init(ds_ready, 8); mbar_init(o_done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmLoc); tma_prefetch_desc(&tmAll);
  }
  pdl_launch_dependents();                             // the dependent (global_norm_bwd_kernel) waits for this grid's completion itself
  GT_MARK(1);
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  GT_MARK(2);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  GT_MARK(3);

  // ---- phase A: logits tile
  gt_issue_logits<kGtStages>(warp, lane, stages, full, empty, s_full, tmem, &tmLoc, &tmAll, row0, col0, ra, ca, KB);
  GT_MARK(4);

  if (warp == 0) {
    // ---- phase C producer: column embeddings (hi, lo) again, 64-wide blocks, read MN-major by the MMA
    if (lane == 0) {
      mbar_wait(ds_ready, 0);                          // dS written => phase-A stages are dead, safe to overwrite
      GT_MARK(5);
      for (int u = u_lo, i = 0; u < u_hi; ++u, ++i) {
        const int slot = i & 1;
        mbar_wait(oempty + slot, ((i >> 1) & 1) ^ 1);
        uint8_t* st = ostg + (size_t)slot * kGtOutStage;
        mbar_expect_tx(ofull + slot, kGtOutStage);
        tma_load_3d(st, &tmAll, ofull + slot, u * 64, col0, ca);
        tma_load_3d(st + kGtTileB, &tmAll, ofull + slot, u * 64, col0, ca + 1);
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    mbar_wait(ds_ready, 0);
    tc_fence_after();
    GT_MARK(5);
    const uint32_t idesc = make_idesc_bf16(kGtM, 64, false, true);       // A = dS K-major (interleaved), B = tile MN-major
    const uint64_t sw0 = make_smem_desc(0, 16, 1024, kLayoutSw128);
    const uint32_t il_lbo = kGtM * 16;
    const uint64_t a_hi = make_smem_desc(smem_u32(dShi), il_lbo, 128, kLayoutNone);
    const uint64_t a_lo = make_smem_desc(smem_u32(dSlo), il_lbo, 128, kLayoutNone);
    const uint32_t a_ks = (2 * il_lbo) >> 4;
    for (int u = u_lo, i = 0; u < u_hi; ++u, ++i) {
      const int slot = i & 1;
      mbar_wait(ofull + slot, (i >> 1) & 1);
      tc_fence_after();
      const uint32_t s0 = smem_u32(ostg + (size_t)slot * kGtOutStage);
      const uint64_t bhi = sw0 | (s0 >> 4), blo = sw0 | ((s0 + kGtTileB) >> 4);
      const uint32_t d = tmem + (uint32_t)(u * 64);                   // D <= 512: the whole dA_hat row block fits TMEM
#pragma unroll
      for (int ks = 0; ks < kGtN / 16; ++ks) {
        umma_ss_w(leader, d, a_hi + ks * a_ks, bhi + ks * 128, idesc, ks != 0);
        umma_ss_w(leader, d, a_hi + ks * a_ks, blo + ks * 128, idesc, true);
        umma_ss_w(leader, d, a_lo + ks * a_ks, bhi + ks * 128, idesc, true);
      }
      umma_commit_w(leader, oempty + slot);
    }
    umma_commit_w(leader, o_done);
    GT_MARK(6);
  } else {
    const int q = warp & 3, h = (warp - 2) >> 2, row = 32 * q + lane, grow = row0 + row;
    const uint32_t trow = tmem + ((uint32_t)(32 * q) << 16);
    float c_self, c_other;
    if (p.cs.on) {
      float c4[4];
      coef_from_src(p.cs, c4);
      c_self = c4[dir]; c_other = c4[1 - dir];
    } else {
      c_self = p.coef[dir]; c_other = p.coef[1 - dir];
    }
    const float lse_self = (grow < p.B) ? p.lse_loc[(size_t)dir * p.B + grow] : 0.f;
    const float* lse_other = p.lse_all + (p.lse_rank_rows ? (size_t)(1 - dir) * p.lse_rank_rows : (size_t)(1 - dir) * p.Bg);
    if (h == 0) {
      // one coalesced load per thread now (overlaps phase A) instead of 128 dependent global loads inside the dS loop
      const int gcol = col0 + row;
      const int gidx = p.lse_rank_rows ? (gcol / p.lse_rank_rows) * p.lse_rank_stride + gcol % p.lse_rank_rows : gcol;
      lse_s[row] = (gcol < p.Bg) ? __ldg(lse_other + gidx) : 0.f;
    }
    // ---- phase B: dS (hi/lo) -> smem (A operand, K = 128 columns of this tile); this warp owns columns [64 h, 64 h + 64)
    mbar_wait(s_full, 0);
    tc_fence_after();
    GT_MARK(10);
    float xs[2][32];
#pragma unroll
    for (int c = 0; c < 2; ++c) tmem_ld32(trow + 64 * h + 32 * c, xs[c]);
    tmem_ld_wait();
    tc_fence_before();
    GT_MARK(11);
    asm volatile("bar.sync 1, 256;" ::: "memory");      // every epilogue warp has its logits in registers: stages may be reused
    GT_MARK(12);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int cb = 64 * h + 32 * c;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int gcol = col0 + cb + j;
        const bool on = grow < p.B && gcol < p.Bg;
        const float s = xs[c][j] * p.scale;
        const float lo_ = lse_s[cb + j];
        float g = c_self * __expf(fminf(s - lse_self, 0.f)) + c_other * __expf(fminf(s - lo_, 0.f));
        g -= (gcol == p.col_offset + grow) ? (c_self + c_other) : 0.f;
        xs[c][j] = on ? g : 0.f;
      }
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        uint4 hi, lo;
        gt_split8(&xs[c][8 * g8], hi, lo);
        const uint32_t off = il_offset(kGtM, row, cb + 8 * g8);
        *reinterpret_cast<uint4*>(dShi + off) = hi;
        *reinterpret_cast<uint4*>(dSlo + off) = lo;
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(ds_ready);
    GT_MARK(13);
    // ---- phase C epilogue: dA_hat partial (this column tile's contribution) -> global; this warp owns half of the D blocks
    mbar_wait(o_done, 0);
    tc_fence_after();
    GT_MARK(14);
    // TMEM -> registers -> per-warp smem transpose tile -> global: every store instruction writes one 128-byte row
    // segment (lane = column) instead of 32 scattered ones (lane = row).  tcgen05.ld is warp-collective: all lanes load.
    float* tile = reinterpret_cast<float*>(ostg) + (warp - 2) * (32 * 36);       // phase-C operand stages are dead now
    const int rr = lane >> 3, c4 = (lane & 7) * 4;
    for (int c0 = 64 * u_lo + 32 * h; c0 < 64 * u_hi; c0 += 64) {
      float x[32];
      tmem_ld32(trow + c0, x);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(tile + lane * 36 + j) = make_float4(x[j] * p.scale, x[j + 1] * p.scale, x[j + 2] * p.scale, x[j + 3] * p.scale);
      __syncwarp();
#pragma unroll
      for (int it = 0; it < 8; ++it) {                   // one instruction = 4 rows x 128 contiguous bytes
        const int r = it * 4 + rr, gr = row0 + 32 * q + r;
        const float4 v4 = *reinterpret_cast<const float4*>(tile + r * 36 + c4);
        if (gr < p.B) {
          if (p.dpart_atomic) {
            // many column tiles (gathered problems): 64 per-tile partials of [B][D] would be 268 MB written and re-read at
            // B = 1024, Bg = 8192 — accumulate in one L2-resident buffer instead (fp32 reductions, order not fixed)
            float* dst = p.dpart + ((size_t)dir * p.B + gr) * D + c0 + c4;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v4.x), "f"(v4.y), "f"(v4.z), "f"(v4.w)
                         : "memory");
          } else {
            *reinterpret_cast<float4*>(p.dpart + (((size_t)dir * nct + ct) * p.B + gr) * D + c0 + c4) = v4;
          }
        }
      }
      __syncwarp();
    }
    tc_fence_before();
    GT_MARK(15);
  }
  __syncthreads();
  GT_MARK(20);
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------
bool global_tc_supported(int B, int Bg, int D) { return D % 64 == 0 && D >= 64 && D <= 512 && B >= 1 && Bg >= B; }

size_t global_tc_workspace_bytes(int B, int Bg, int D) {
  const int nct = (Bg + kGtN - 1) / kGtN;
  const size_t split = (size_t)4 * Bg * D * sizeof(bf16) + (size_t)4 * B * D * sizeof(bf16);   // all + local, hi/lo x a/b
  const size_t fwd = ((size_t)8 * nct * B + 2 * (size_t)B) * sizeof(float);      // 2 partials per column tile (one per half)
  const size_t bwd = (size_t)2 * nct * B * D * sizeof(float);
  return ((split + 255) & ~(size_t)255) + (fwd > bwd ? fwd : bwd);
}

struct GtHost {
  bf16* all_split; bf16* loc_split; float* scratch; int nct;
};
static GtHost gt_carve(void* ws, int B, int Bg, int D) {
  GtHost h;
  h.nct = (Bg + kGtN - 1) / kGtN;
  h.all_split = (bf16*)ws;
  h.loc_split = h.all_split + (size_t)4 * Bg * D;
  const size_t split = (size_t)4 * Bg * D * sizeof(bf16) + (size_t)4 * B * D * sizeof(bf16);
  h.scratch = (float*)((uint8_t*)ws + ((split + 255) & ~(size_t)255));
  return h;
}

static int gt_maps(const GtHost& h, int B, int Bg, int D, int col_offset, CUtensorMap* tmLoc, CUtensorMap* tmAll) {
  // [which (a_hi, a_lo, b_hi, b_lo)][row][d]: the `which` index is the 3rd TMA coordinate.  The local rows are the window
  // [col_offset, col_offset + B) of the all-rows array (rows beyond B are out of bounds of the window: zero fill).
  int rc;
  if ((rc = make_tmap_bf16_3d(tmLoc, h.all_split + (size_t)col_offset * D, D, B, 4, 64, kGtM, false, (uint64_t)Bg * D)) != CFA_OK) return rc;
  if ((rc = make_tmap_bf16_3d(tmAll, h.all_split, D, Bg, 4, 64, kGtN)) != CFA_OK) return rc;
  return CFA_OK;
}

int global_tc_fwd(const float* a_loc, const float* b_loc, const float* a_all, const float* b_all, int B, int Bg, int D,
                  int col_offset, float scale, float eps, float* norms2, float** part_m, float** part_l, float** diag,
                  int* nsplit, void* ws, int gathered_ranks, const PeerTable* peers, cudaStream_t st) {
  const GtHost h = gt_carve(ws, B, Bg, D);
  PeerTable none{};
  // ONE launch: every global row is normalised and split once; the local rows' norms come out of the same pass
  if (peers && peers->n > 1)
    CFA_CUDA_TRY(cfa_launch_pdl(2, gt_split_kernel, dim3((Bg + 7) / 8, 2), dim3(256), 0, st, (const float*)nullptr, (const float*)nullptr, Bg, D, eps,
                                h.all_split, norms2, B, (size_t)0, *peers, col_offset, B));
  else if (gathered_ranks > 1)
    CFA_CUDA_TRY(cfa_launch_pdl(2, gt_split_kernel, dim3((Bg + 7) / 8, 2), dim3(256), 0, st, a_all, b_all, Bg, D, eps, h.all_split, norms2, B,
                                (size_t)2 * B * D, none, col_offset, B));
  else
    CFA_CUDA_TRY(cfa_launch_pdl(2, gt_split_kernel, dim3((Bg + 7) / 8, 2), dim3(256), 0, st, a_all, b_all, Bg, D, eps, h.all_split, norms2, Bg,
                                (size_t)0, none, col_offset, B));
  CFA_CUDA_TRY(cudaGetLastError());
  CUtensorMap tmLoc, tmAll;
  int rc = gt_maps(h, B, Bg, D, col_offset, &tmLoc, &tmAll);
  if (rc != CFA_OK) return rc;
  GtParams p{};
  p.B = B; p.Bg = Bg; p.D = D; p.col_offset = col_offset; p.scale = scale;
  p.part_m = h.scratch; p.part_l = p.part_m + (size_t)4 * h.nct * B; p.diag = p.part_l + (size_t)4 * h.nct * B;
  // small grids (under two CTAs per SM): one CTA per SM with a 3-deep ring; large grids: single-stage CTAs, three per SM
  const dim3 grid((B + kGtM - 1) / kGtM, h.nct, 2);
  const bool big = (size_t)grid.x * grid.y * grid.z >= 2 * 148;
  const size_t smem = (big ? kGtFwdStages : kGtStages) * kGtStage + 1024 + 1024;
  CFA_SMEM_ATTR_ONCE(gt_fwd_kernel<kGtFwdStages>, kGtFwdStages * kGtStage + 2048);
  CFA_SMEM_ATTR_ONCE(gt_fwd_kernel<kGtStages>, kGtStages * kGtStage + 2048);
  if (big) CFA_CUDA_TRY(cfa_launch_pdl(2, gt_fwd_kernel<kGtFwdStages>, grid, dim3(kGtThreads), smem, st, tmLoc, tmAll, p));
  else CFA_CUDA_TRY(cfa_launch_pdl(2, gt_fwd_kernel<kGtStages>, grid, dim3(kGtThreads), smem, st, tmLoc, tmAll, p));
  *part_m = p.part_m; *part_l = p.part_l; *diag = p.diag; *nsplit = 2 * h.nct;
  return launch_status();
}

int global_tc_bwd(int B, int Bg, int D, int col_offset, float scale, const float* lse_loc2, const float* lse_all2,
                  const float* coef2, float** dpart, int* nsplit, void* ws, int gathered_ranks, cudaStream_t st) {
  // the split operands written by the forward are still in the workspace
  const GtHost h = gt_carve(ws, B, Bg, D);
  CUtensorMap tmLoc, tmAll;
  int rc = gt_maps(h, B, Bg, D, col_offset, &tmLoc, &tmAll);
  if (rc != CFA_OK) return rc;
  GtParams p{};
  p.B = B; p.Bg = Bg; p.D = D; p.col_offset = col_offset; p.scale = scale;
  p.lse_loc = lse_loc2; p.lse_all = lse_all2; p.coef = coef2; p.dpart = h.scratch; p.dbg = g_gt_dbg;
  if (g_coef_src) p.cs = *g_coef_src;
  p.dpart_atomic = h.nct > 2;
  if (p.dpart_atomic) CFA_CUDA_TRY(cudaMemsetAsync(h.scratch, 0, (size_t)2 * B * D * sizeof(float), st));
  p.lse_rank_rows = gathered_ranks > 1 ? B : 0; p.lse_rank_stride = gathered_ranks > 1 ? 2 * B + 2 : 0;
  const size_t smem = kGtStages * kGtStage + 1024 + 1024;
  CFA_SMEM_ATTR_ONCE(gt_bwd_kernel, smem);
  // few tiles (rank-local batches): divide the D / 64 output blocks over up to 8 CTAs per tile so that the grid covers
  // the GPU -- each CTA recomputes the logits tile (cheap) and contracts / writes only its blocks
  const int tiles = ((B + kGtM - 1) / kGtM) * h.nct * 2, KB = D / 64;
  int ds = 1;
  while (ds < 8 && KB % (2 * ds) == 0 && tiles * 2 * ds <= 148) ds *= 2;
  p.dsplit = ds;
  gt_bwd_kernel<<<dim3((B + kGtM - 1) / kGtM, h.nct, 2 * ds), kGtThreads, smem, st>>>(tmLoc, tmAll, p);
  *dpart = h.scratch; *nsplit = p.dpart_atomic ? 1 : h.nct;
  return launch_status();
}

}  // namespace cfa

// tuning / debugging aid: host-mapped int buffer receiving per-warp progress markers of gt_bwd_kernel
extern "C" int cfa_debug_set_marker_buffer(void* mapped_buffer) {
  cfa::g_gt_dbg = (int*)mapped_buffer;
  return 0;
}
