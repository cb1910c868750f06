// Global InfoNCE for the rank-local case (Bg == B, every row is also a column): fp32-exact CUDA-core tiles built for
// LATENCY.  At B = 256 the whole problem is 67 MFLOP, so the only thing that matters is how many short, independent
// CTAs there are and how few dependent steps each one takes.
//
// One 32 x 32 tile of the logits S = scale * a^_i . b^_j serves BOTH directions (losses.py:215-216 calls
// pairwise_contrastive_loss twice with swapped arguments, i.e. on S and on S^T):
//   forward : row-wise (max, sum exp) partials -> direction a, column-wise partials -> direction b, diagonal -> targets;
//   backward: dS = c_a softmax_row + c_b softmax_col - (c_a + c_b) I once, then
//             da_part = scale * (dS / |b|) . b_raw   and   db_part = scale * (dS^T / |a|) . a_raw   for one D slice.
// Row norms (F.normalize, losses.py:152-153) are accumulated while the K chunks stream through shared memory.
// Partials are merged by global_combine_kernel / global_norm_bwd_kernel (global_infonce.cu).
#include "common.cuh"
#include "global_combine.cuh"
#include <math_constants.h>
#include <cstdlib>

namespace cfa {

// global_sym_fwd_kernel draws tickets from a counter that lives in the CALLER's workspace (zeroed by a memset node in
// front of the launch): the CTA that draws the last ticket merges the partials (global_combine_body), so no separate
// single-CTA launch is needed and concurrent forwards on different streams do not interfere.
struct SymCombine {
  float* lse; float* sums2; const float* local_partial; const uint8_t* mask; int T; float gw, lw; float* out8;
  unsigned int* ticket;
};

constexpr int kSyT = 32;             // tile rows = tile cols
constexpr int kSyK = 128;            // K chunk
constexpr int kSyLd = kSyK + 4;      // row stride (floats): float4 reads of 8 consecutive rows hit 8 distinct bank groups
constexpr int kSyThreads = 256;

// One K chunk of 32 rows of `a` and 32 rows of `b` -> smem (warp w loads rows w, w+8, w+16, w+24: one float4 per lane).
// ALL eight global loads are issued before the first use (the shuffles of the norm reduction would otherwise serialise
// them: 8 dependent L2 round trips per chunk).  kNorm: also accumulate the squared row norms (every lane gets the sum).
__device__ __forceinline__ float4 sy_ld4(const float* __restrict__ src, int row, int rows, int D, int d) {
  float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row < rows && d < D) {
    const float* s = src + (size_t)row * D;
    if (d + 3 < D && (D & 3) == 0) x = __ldg(reinterpret_cast<const float4*>(s + d));
    else { x.x = s[d]; x.y = d + 1 < D ? s[d + 1] : 0.f; x.z = d + 2 < D ? s[d + 2] : 0.f; x.w = d + 3 < D ? s[d + 3] : 0.f; }
  }
  return x;
}
// global -> registers (all eight loads in flight), registers -> smem (+ optional squared row norms) as separate steps so
// that the loads of chunk k+1 can be issued before the arithmetic of chunk k (register double buffering)
struct SyRegs { float4 a[4], b[4]; };
__device__ __forceinline__ void sy_fetch(SyRegs& x, const float* __restrict__ a, const float* __restrict__ b, int r0, int c0,
                                         int rows, int D, int d0, int warp, int lane) {
  const int d = d0 + 4 * lane;
#pragma unroll
  for (int i = 0; i < 4; ++i) { x.a[i] = sy_ld4(a, r0 + warp + 8 * i, rows, D, d); x.b[i] = sy_ld4(b, c0 + warp + 8 * i, rows, D, d); }
}
template <bool kNorm>
__device__ __forceinline__ void sy_commit(const SyRegs& x, float* __restrict__ As, float* __restrict__ Bs, float* ssa, float* ssb,
                                          int warp, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    *reinterpret_cast<float4*>(As + (warp + 8 * i) * kSyLd + 4 * lane) = x.a[i];
    *reinterpret_cast<float4*>(Bs + (warp + 8 * i) * kSyLd + 4 * lane) = x.b[i];
  }
  if (kNorm) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ssa[i] += warp_sum(fmaf(x.a[i].x, x.a[i].x, fmaf(x.a[i].y, x.a[i].y, fmaf(x.a[i].z, x.a[i].z, x.a[i].w * x.a[i].w))));
      ssb[i] += warp_sum(fmaf(x.b[i].x, x.b[i].x, fmaf(x.b[i].y, x.b[i].y, fmaf(x.b[i].z, x.b[i].z, x.b[i].w * x.b[i].w))));
    }
  }
}

// acc[i][j] += sum_k A[ty + 16 i][k] * B[tx + 16 j][k] over one chunk
__device__ __forceinline__ void sy_mac_chunk(const float* __restrict__ As, const float* __restrict__ Bs, float (&acc)[2][2],
                                             int ty, int tx) {
#pragma unroll 8
  for (int k = 0; k < kSyK; k += 4) {
    const float4 a0 = *reinterpret_cast<const float4*>(As + ty * kSyLd + k);
    const float4 a1 = *reinterpret_cast<const float4*>(As + (ty + 16) * kSyLd + k);
    const float4 b0 = *reinterpret_cast<const float4*>(Bs + tx * kSyLd + k);
    const float4 b1 = *reinterpret_cast<const float4*>(Bs + (tx + 16) * kSyLd + k);
    acc[0][0] = fmaf(a0.x, b0.x, fmaf(a0.y, b0.y, fmaf(a0.z, b0.z, fmaf(a0.w, b0.w, acc[0][0]))));
    acc[0][1] = fmaf(a0.x, b1.x, fmaf(a0.y, b1.y, fmaf(a0.z, b1.z, fmaf(a0.w, b1.w, acc[0][1]))));
    acc[1][0] = fmaf(a1.x, b0.x, fmaf(a1.y, b0.y, fmaf(a1.z, b0.z, fmaf(a1.w, b0.w, acc[1][0]))));
    acc[1][1] = fmaf(a1.x, b1.x, fmaf(a1.y, b1.y, fmaf(a1.z, b1.z, fmaf(a1.w, b1.w, acc[1][1]))));
  }
}

// raw logits tile (K loop over D) + clamped norms of the tile's rows (nrm[0..31]) and columns (nrm[32..63])
__device__ __forceinline__ void sy_logits_tile(const float* __restrict__ a, const float* __restrict__ b, int B, int D, int r0,
                                               int c0, float eps, float* As, float* Bs, float* nrm, float (&acc)[2][2]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float ssa[4] = {0.f, 0.f, 0.f, 0.f}, ssb[4] = {0.f, 0.f, 0.f, 0.f};
  acc[0][0] = acc[0][1] = acc[1][0] = acc[1][1] = 0.f;
  SyRegs cur, nxt;
  sy_fetch(cur, a, b, r0, c0, B, D, 0, warp, lane);
  for (int d0 = 0; d0 < D; d0 += kSyK) {
    __syncthreads();
    sy_commit<true>(cur, As, Bs, ssa, ssb, warp, lane);
    if (d0 + kSyK < D) sy_fetch(nxt, a, b, r0, c0, B, D, d0 + kSyK, warp, lane);     // in flight during the MACs below
    __syncthreads();
    sy_mac_chunk(As, Bs, acc, ty, tx);
    cur = nxt;
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      nrm[warp + 8 * i] = fmaxf(sqrtf(ssa[i]), eps);
      nrm[32 + warp + 8 * i] = fmaxf(sqrtf(ssb[i]), eps);
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------------
// forward: grid (row tiles, column tiles)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSyThreads)
global_sym_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int B, int D, float scale, float eps,
                      float* part_m, float* part_l, float* diag, float* __restrict__ norms /* [2][B] */, const SymCombine cb) {
  __shared__ __align__(16) float As[kSyT * kSyLd];
  __shared__ __align__(16) float Bs[kSyT * kSyLd];
  __shared__ float nrm[64];
  __shared__ float St[kSyT][kSyT + 1];
  const int rt = blockIdx.x, ct = blockIdx.y, nt = gridDim.x;     // square problem: gridDim.x == gridDim.y
  const int r0 = rt * kSyT, c0 = ct * kSyT;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[2][2];
  sy_logits_tile(a, b, B, D, r0, c0, eps, As, Bs, nrm, acc);
  if (ct == 0 && threadIdx.x < kSyT && r0 + threadIdx.x < B) norms[r0 + threadIdx.x] = nrm[threadIdx.x];
  if (rt == 0 && threadIdx.x < kSyT && c0 + threadIdx.x < B) norms[(size_t)B + c0 + threadIdx.x] = nrm[32 + threadIdx.x];
  // scaled logits; out-of-range entries are -inf
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = ty + 16 * i, grow = r0 + m;
    float v[2], tmax = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int n = tx + 16 * j, gcol = c0 + n;
      v[j] = (grow < B && gcol < B) ? scale * (acc[i][j] / (nrm[m] * nrm[32 + n])) : -CUDART_INF_F;
      St[m][n] = v[j];
      tmax = fmaxf(tmax, v[j]);
      if (grow < B && gcol == grow) { diag[grow] = v[j]; diag[(size_t)B + grow] = v[j]; }
    }
    // direction a: this row against the 32 columns of the tile (16 lanes share a row)
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 2; ++j) s += (v[j] == -CUDART_INF_F) ? 0.f : expf(v[j] - tmax);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (tx == 0 && grow < B) {
      part_m[((size_t)0 * nt + ct) * B + grow] = tmax;
      part_l[((size_t)0 * nt + ct) * B + grow] = s;
    }
  }
  __syncthreads();
  {
    // direction b: column n against the 32 rows of the tile (8 lanes share a column, 4 rows each)
    const int n = threadIdx.x >> 3, part = threadIdx.x & 7, gcol = c0 + n;
    float v[4], tmax = -CUDART_INF_F;
#pragma unroll
    for (int k = 0; k < 4; ++k) { v[k] = St[part * 4 + k][n]; tmax = fmaxf(tmax, v[k]); }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += (v[k] == -CUDART_INF_F) ? 0.f : expf(v[k] - tmax);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (part == 0 && gcol < B) {
      part_m[((size_t)1 * nt + rt) * B + gcol] = tmax;
      part_l[((size_t)1 * nt + rt) * B + gcol] = s;
    }
  }
  // last CTA: merge the partials of the whole grid (classic threadfence-reduction hand-off)
  __shared__ unsigned int s_last;
  __shared__ float red[8];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    const unsigned int t = atomicAdd(cb.ticket, 1u);
    s_last = (t == total - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    global_combine_body(part_m, part_l, diag, B, nt, cb.lse, cb.sums2, B, cb.local_partial, cb.mask, cb.T, cb.gw, cb.lw, cb.out8, red);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// backward: grid (row tiles, column tiles, D slices of kSyDz)
// ------------------------------------------------------------------------------------------------------------------
constexpr int kSyDz = 256;           // output D slice per CTA (two K chunks)

__global__ void __launch_bounds__(kSyThreads)
global_sym_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int B, int D, float scale, float eps,
                      const float* __restrict__ lse2 /* [2][B] */, const float* __restrict__ coef /* c_a, c_b */,
                      float* __restrict__ out /* [2][nt][B][D] */) {
  __shared__ __align__(16) float As[kSyT * kSyLd];
  __shared__ __align__(16) float Bs[kSyT * kSyLd];
  __shared__ float nrm[64];
  __shared__ float dSa[kSyT][kSyT + 1];      // dS_ij / |b_j|   (for da)
  __shared__ float dSb[kSyT][kSyT + 1];      // dS_ij / |a_i|   (for db), stored transposed: [j][i]
  const int rt = blockIdx.x, ct = blockIdx.y, nt = gridDim.x;
  const int r0 = rt * kSyT, c0 = ct * kSyT, dz0 = blockIdx.z * kSyDz;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const float c_a = coef[0], c_b = coef[1];
  float acc[2][2];
  sy_logits_tile(a, b, B, D, r0, c0, eps, As, Bs, nrm, acc);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int m = ty + 16 * i, n = tx + 16 * j, grow = r0 + m, gcol = c0 + n;
      float ga = 0.f, gb = 0.f;                    // out-of-range rows / columns have zero norms (eps may be 0): keep them 0
      if (grow < B && gcol < B) {
        const float s = scale * (acc[i][j] / (nrm[m] * nrm[32 + n]));
        float g = c_a * expf(s - lse2[grow]) + c_b * expf(s - lse2[(size_t)B + gcol]);
        if (gcol == grow) g -= (c_a + c_b);
        ga = g / nrm[32 + n];
        gb = g / nrm[m];
      }
      dSa[m][n] = ga;
      dSb[n][m] = gb;
    }
  // thread -> output row r (of the 32) and columns cg + 32 c4 + {0..3}, c4 < 4, of a 128-wide chunk: the 8 lanes that
  // share a row read / write 128 contiguous bytes per access (conflict-free float4 smem reads, coalesced stores)
  const int r = threadIdx.x >> 3, cg = (threadIdx.x & 7) * 4;
  float* ssd = nullptr;
  const int dz1 = min(dz0 + kSyDz, D);
  SyRegs cur, nxt;
  sy_fetch(cur, a, b, r0, c0, B, D, dz0, warp, lane);
  for (int d0 = dz0; d0 < dz1; d0 += kSyK) {
    __syncthreads();
    sy_commit<false>(cur, As, Bs, ssd, ssd, warp, lane);
    if (d0 + kSyK < dz1) sy_fetch(nxt, a, b, r0, c0, B, D, d0 + kSyK, warp, lane);
    __syncthreads();
    float oa[16], ob[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) { oa[c] = 0.f; ob[c] = 0.f; }
#pragma unroll 4
    for (int k = 0; k < kSyT; ++k) {
      const float ga = dSa[r][k], gb = dSb[r][k];
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float4 bv = *reinterpret_cast<const float4*>(Bs + k * kSyLd + cg + 32 * c4);
        const float4 av = *reinterpret_cast<const float4*>(As + k * kSyLd + cg + 32 * c4);
        oa[4 * c4] = fmaf(ga, bv.x, oa[4 * c4]); oa[4 * c4 + 1] = fmaf(ga, bv.y, oa[4 * c4 + 1]);
        oa[4 * c4 + 2] = fmaf(ga, bv.z, oa[4 * c4 + 2]); oa[4 * c4 + 3] = fmaf(ga, bv.w, oa[4 * c4 + 3]);
        ob[4 * c4] = fmaf(gb, av.x, ob[4 * c4]); ob[4 * c4 + 1] = fmaf(gb, av.y, ob[4 * c4 + 1]);
        ob[4 * c4 + 2] = fmaf(gb, av.z, ob[4 * c4 + 2]); ob[4 * c4 + 3] = fmaf(gb, av.w, ob[4 * c4 + 3]);
      }
    }
    const int d = d0 + cg;
    const bool vec = (D & 3) == 0;
    if (r0 + r < B) {
      float* dst = out + (((size_t)0 * nt + ct) * B + r0 + r) * D + d;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        if (vec && d + 32 * c4 + 3 < D)
          *reinterpret_cast<float4*>(dst + 32 * c4) = make_float4(oa[4 * c4] * scale, oa[4 * c4 + 1] * scale, oa[4 * c4 + 2] * scale, oa[4 * c4 + 3] * scale);
        else
          for (int e = 0; e < 4; ++e) if (d + 32 * c4 + e < D) dst[32 * c4 + e] = oa[4 * c4 + e] * scale;
      }
    }
    if (c0 + r < B) {
      float* dst = out + (((size_t)1 * nt + rt) * B + c0 + r) * D + d;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        if (vec && d + 32 * c4 + 3 < D)
          *reinterpret_cast<float4*>(dst + 32 * c4) = make_float4(ob[4 * c4] * scale, ob[4 * c4 + 1] * scale, ob[4 * c4 + 2] * scale, ob[4 * c4 + 3] * scale);
        else
          for (int e = 0; e < 4; ++e) if (d + 32 * c4 + e < D) dst[32 * c4 + e] = ob[4 * c4 + e] * scale;
      }
    }
    cur = nxt;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------------------------
bool global_sym_supported(int B, int Bg, int D) { return Bg == B && B >= 1 && B <= 512 && D >= 1; }
int global_sym_tiles(int B) { return (B + kSyT - 1) / kSyT; }

size_t global_sym_workspace_bytes(int B, int D) {
  const size_t nt = global_sym_tiles(B);
  const size_t fwd = ((size_t)4 * nt * B + 2 * (size_t)B) * sizeof(float) + 128;      // + ticket counter
  const size_t bwd = (size_t)2 * nt * B * D * sizeof(float);
  return fwd > bwd ? fwd : bwd;
}

// forward including the merge of the partials (and the optional SPARC scalar epilogue): ONE launch
int global_sym_fwd(const float* a, const float* b, int B, int D, float scale, float eps, float* norms2, float* lse2,
                   float* sums2, const float* local_partial, const uint8_t* mask, int T, float gw, float lw, float* out8,
                   void* ws, cudaStream_t st) {
  const int nt = global_sym_tiles(B);
  float* part_m = (float*)ws;
  float* part_l = part_m + (size_t)2 * nt * B;
  float* diag = part_l + (size_t)2 * nt * B;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(diag + (((size_t)2 * B + 31) & ~(size_t)31));
  CFA_CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned int), st));
  const SymCombine cb{lse2, sums2, local_partial, mask, T, gw, lw, out8, ticket};
  global_sym_fwd_kernel<<<dim3(nt, nt), kSyThreads, 0, st>>>(a, b, B, D, scale, eps, part_m, part_l, diag, norms2, cb);
  return launch_status();
}

int global_sym_bwd(const float* a, const float* b, int B, int D, float scale, float eps, const float* lse2, const float* coef2,
                   float** dpart, int* nsplit, void* ws, cudaStream_t st) {
  const int nt = global_sym_tiles(B);
  *dpart = (float*)ws;
  *nsplit = nt;
  global_sym_bwd_kernel<<<dim3(nt, nt, (D + kSyDz - 1) / kSyDz), kSyThreads, 0, st>>>(a, b, B, D, scale, eps, lse2, coef2,
                                                                                      (float*)ws);
  return launch_status();
}

}  // namespace cfa
