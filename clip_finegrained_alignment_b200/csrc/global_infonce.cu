// Global (batch-level) InfoNCE, both directions per launch, fp32 CUDA-core tiles (losses.py:14-36, 145-163, 207-217).
//
// Inputs are the RAW pooled embeddings: local rows a_loc/b_loc [B,D] and the (all-gathered) columns a_all/b_all
// [Bg,D].  L2 normalisation (F.normalize, eps) is folded into the kernels: row / column norms are accumulated
// while the K chunks stream through shared memory and the logits tile is scaled afterwards, so there is no
// separate normalise pass and the B x Bg logits never reach memory.
//   direction 0: rows = a_loc, columns = b_all  (image -> text);   direction 1: rows = b_loc, columns = a_all.
#include "common.cuh"
#include <cstdlib>
#include "simt_tile.cuh"
#include "global_combine.cuh"
#include "sparc_paths.h"
#include <math_constants.h>

namespace cfa {

constexpr int kGR = 32, kGC = 64, kGK = 128, kGLd = kGK + 1;    // forward tile: 32 rows x 64 cols, K chunks of 128 (few, long loads)

__global__ void __launch_bounds__(kNT)
global_fwd_kernel(const float* __restrict__ a_loc, const float* __restrict__ b_loc, const float* __restrict__ a_all,
                  const float* __restrict__ b_all, int B, int Bg, int D, int col_offset, float scale, float eps,
                  float* __restrict__ part_m, float* __restrict__ part_l, float* __restrict__ diag,
                  float* __restrict__ norms /* [2][B] clamped row norms */) {
  extern __shared__ float gsm[];
  float* stA = gsm;                        // [32 x 129]
  float* stB = stA + kGR * kGLd;           // [64 x 129]
  float* nrm = stB + kGC * kGLd;           // [96]
  const int dir = blockIdx.z, split = blockIdx.y, nsplit = gridDim.y, r0 = blockIdx.x * kGR;
  const float* rows = dir ? b_loc : a_loc;
  const float* cols = dir ? a_all : b_all;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int rows_valid = min(kGR, B - r0);
  float run_m[2], run_l[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) { run_m[i] = -CUDART_INF_F; run_l[i] = 0.f; }
  const int ntiles = (Bg + kGC - 1) / kGC;
  for (int ct = split; ct < ntiles; ct += nsplit) {
    const int c0 = ct * kGC;
    const int cols_valid = min(kGC, Bg - c0);
    float acc[2][4];
    tile_zero(acc);
    float ss = 0.f;                                   // thread t < 96 owns the squared norm of one row / column
    for (int d0 = 0; d0 < D; d0 += kGK) {
      __syncthreads();
      load_tile<float>(stA, kGLd, rows + (size_t)r0 * D, kGR, rows_valid, D, d0, kGK);
      load_tile<float>(stB, kGLd, cols + (size_t)c0 * D, kGC, cols_valid, D, d0, kGK);
      __syncthreads();
      if (threadIdx.x < kGR + kGC) {
        const float* src = threadIdx.x < kGR ? stA + threadIdx.x * kGLd : stB + (threadIdx.x - kGR) * kGLd;
#pragma unroll 8
        for (int k = 0; k < kGK; ++k) ss = fmaf(src[k], src[k], ss);
      }
      tile_mac<2, 4>(acc, 0, 0, kGR, kGC, kGK, [&](int m, int k) { return stA[m * kGLd + k]; },
                     [&](int k, int n) { return stB[n * kGLd + k]; });
    }
    __syncthreads();
    if (threadIdx.x < kGR + kGC) nrm[threadIdx.x] = fmaxf(sqrtf(ss), eps);
    __syncthreads();
    if (ct == split && split == 0 && threadIdx.x < rows_valid) norms[(size_t)dir * B + r0 + threadIdx.x] = nrm[threadIdx.x];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int m = ty + 16 * i, grow = r0 + m;
      float v[4], tmax = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = tx + 16 * j, gcol = c0 + n;
        v[j] = (gcol < Bg) ? scale * (acc[i][j] / (nrm[m] * nrm[kGR + n])) : -CUDART_INF_F;
        tmax = fmaxf(tmax, v[j]);
        if (grow < B && gcol == col_offset + grow) diag[(size_t)dir * B + grow] = v[j];
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
      const float new_m = fmaxf(run_m[i], tmax);
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) s += (v[j] == -CUDART_INF_F) ? 0.f : expf(v[j] - new_m);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      run_l[i] = run_l[i] * ((run_m[i] == -CUDART_INF_F) ? 0.f : expf(run_m[i] - new_m)) + s;
      run_m[i] = new_m;
    }
  }
  if (tx == 0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int grow = r0 + ty + 16 * i;
      if (grow < B) {
        part_m[((size_t)dir * nsplit + split) * B + grow] = run_m[i];
        part_l[((size_t)dir * nsplit + split) * B + grow] = run_l[i];
      }
    }
  }
}

__global__ void __launch_bounds__(kNT)
global_combine_kernel(const float* __restrict__ part_m, const float* __restrict__ part_l, const float* __restrict__ diag,
                      int B, int nsplit, float* __restrict__ lse, float* __restrict__ sums2, int global_batch,
                      const float* __restrict__ local_partial, const uint8_t* __restrict__ mask, int T, float gw, float lw,
                      float* __restrict__ out8) {
  __shared__ float red[8];
  pdl_launch_dependents();                             // programmatic dependent launch (common.cuh): start early, wait here
  pdl_wait();
  global_combine_body(part_m, part_l, diag, B, nsplit, lse, sums2, global_batch, local_partial, mask, T, gw, lw, out8, red);
}

// Many column splits (gathered problems: Bg / 128 tiles per row): merge the online-softmax partials with one thread per
// (direction, row) over a whole grid first; the single-CTA kernel then only sums 2B cross-entropies (nsplit = 0).
// block = 32 rows x 8 split groups: loads are coalesced along the rows, every thread merges nsplit / 8 partials and the
// 8 group results of a row are merged through shared memory in a fixed order.
__global__ void __launch_bounds__(256)
global_merge_rows_kernel(const float* __restrict__ part_m, const float* __restrict__ part_l, int B, int nsplit,
                         float* __restrict__ lse) {
  __shared__ float sm[8][32], sl[8][32];
  pdl_launch_dependents();
  pdl_wait();
  const int tx = threadIdx.x & 31, sg = threadIdx.x >> 5;
  const int idx = blockIdx.x * 32 + tx;
  float M = -CUDART_INF_F, Lq = 0.f;
  if (idx < 2 * B) {
    const int dir = idx / B, i = idx - dir * B;
    const float* pm = part_m + (size_t)dir * nsplit * B + i;
    const float* pl = part_l + (size_t)dir * nsplit * B + i;
    for (int s = sg; s < nsplit; s += 8) {
      const float m = pm[(size_t)s * B], l = pl[(size_t)s * B];
      if (m == -CUDART_INF_F) continue;
      const float nm = fmaxf(M, m);
      Lq = Lq * ((M == -CUDART_INF_F) ? 0.f : expf(M - nm)) + l * expf(m - nm);
      M = nm;
    }
  }
  sm[sg][tx] = M; sl[sg][tx] = Lq;
  __syncthreads();
  if (sg == 0 && idx < 2 * B) {
    float Mx = -CUDART_INF_F;
#pragma unroll
    for (int k = 0; k < 8; ++k) Mx = fmaxf(Mx, sm[k][tx]);
    float L = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) if (sm[k][tx] != -CUDART_INF_F) L += sl[k][tx] * expf(sm[k][tx] - Mx);
    lse[idx] = Mx + logf(L);
  }
}

// ------------------------------------------------------------------------------------------------
// backward: per 32 x 32 logits tile recompute S (with on-the-fly norms), form
//   dS_ij = c_self exp(S_ij - lse_self[i]) + c_other exp(S_ij - lse_other[j]) - (c_self + c_other) [j == off + i]
// and accumulate  sum_j (dS_ij / |col_j|) col_j  for a 256-wide slice of D.  Partials per column split.
// ------------------------------------------------------------------------------------------------
constexpr int kBR = 32, kBC = 32, kBK = 128, kBLd = kBK + 1, kBDz = 256, kBLdB = kBDz + 1, kBLdS = 33;

__global__ void __launch_bounds__(kNT)
global_bwd_kernel(const float* __restrict__ a_loc, const float* __restrict__ b_loc, const float* __restrict__ a_all,
                  const float* __restrict__ b_all, int B, int Bg, int D, int col_offset, float scale, float eps,
                  const float* __restrict__ lse_loc /* [2][B] */, const float* __restrict__ lse_all /* [2][Bg] */,
                  const float* __restrict__ norms /* [2][B] */, const float* __restrict__ coef /* c_a, c_b */,
                  float* __restrict__ out /* [2][nsplit][B][D] */) {
  extern __shared__ float smem[];
  float* stA = smem;                       // [32 x 129]
  float* stB = stA + kBR * kBLd;           // [32 x 129]
  float* dS = stB + kBC * kBLd;            // [32 x 33]
  float* bt = dS + kBR * kBLdS;            // [32 x 257]
  float* cn = bt + kBC * kBLdB;            // [32] column norms
  const int rb = (B + kBR - 1) / kBR;
  const int dir = blockIdx.x / rb, r0 = (blockIdx.x - dir * rb) * kBR;
  const int split = blockIdx.y, nsplit = gridDim.y, dz0 = blockIdx.z * kBDz;
  const float* rows = dir ? b_loc : a_loc;
  const float* cols = dir ? a_all : b_all;
  const float* lse_self = lse_loc + (size_t)dir * B;            // rows' own direction
  const float* lse_other = lse_all + (size_t)(1 - dir) * Bg;    // the columns' direction, global
  const float* rn = norms + (size_t)dir * B;
  const float c_self = coef[dir], c_other = coef[1 - dir];
  const int dzn = min(kBDz, D - dz0);
  const int rows_valid = min(kBR, B - r0);
  float oacc[2][16];
  tile_zero(oacc);
  const int ntiles = (Bg + kBC - 1) / kBC;
  for (int ct = split; ct < ntiles; ct += nsplit) {
    const int c0 = ct * kBC;
    const int cols_valid = min(kBC, Bg - c0);
    float acc[2][2];
    tile_zero(acc);
    float ss = 0.f;
    for (int d0 = 0; d0 < D; d0 += kBK) {
      __syncthreads();
      load_tile<float>(stA, kBLd, rows + (size_t)r0 * D, kBR, rows_valid, D, d0, kBK);
      load_tile<float>(stB, kBLd, cols + (size_t)c0 * D, kBC, cols_valid, D, d0, kBK);
      __syncthreads();
      if (threadIdx.x < kBC) {
        const float* src = stB + threadIdx.x * kBLd;
#pragma unroll 8
        for (int k = 0; k < kBK; ++k) ss = fmaf(src[k], src[k], ss);
      }
      tile_mac<2, 2>(acc, 0, 0, kBR, kBC, kBK, [&](int m, int k) { return stA[m * kBLd + k]; },
                     [&](int k, int n) { return stB[n * kBLd + k]; });
    }
    if (threadIdx.x < kBC) cn[threadIdx.x] = fmaxf(sqrtf(ss), eps);
    for (int idx = threadIdx.x; idx < kBC * dzn; idx += kNT) {   // raw column rows for the output contraction
      const int r = idx / dzn, c = idx - r * dzn;
      bt[r * kBLdB + c] = (r < cols_valid) ? cols[(size_t)(c0 + r) * D + dz0 + c] : 0.f;
    }
    __syncthreads();
    {
      const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int m = ty + 16 * i, n = tx + 16 * j;
          const int grow = r0 + m, gcol = c0 + n;
          float g = 0.f;
          if (grow < B && gcol < Bg) {
            const float s = scale * (acc[i][j] / (rn[grow] * cn[n]));
            g = c_self * expf(s - lse_self[grow]) + c_other * expf(s - lse_other[gcol]);
            if (gcol == col_offset + grow) g -= (c_self + c_other);
            g /= cn[n];                                           // d/d(col_hat) -> raw column, first factor
          }
          dS[m * kBLdS + n] = g;
        }
    }
    __syncthreads();
    tile_mac<2, 16>(oacc, 0, 0, kBR, dzn, kBC, [&](int m, int k) { return dS[m * kBLdS + k]; },
                    [&](int k, int n) { return bt[k * kBLdB + n]; });
  }
  tile_foreach<2, 16>(oacc, 0, 0, rows_valid, dzn, [&](int m, int n, float x) {
    out[(((size_t)dir * nsplit + split) * B + r0 + m) * D + dz0 + n] = x * scale;
  });
}

// d(raw row) = J_n^T (sum of partials):  (g - x_hat (x_hat . g)) / max(|x|, eps)      one CTA per (row, direction)
__global__ void __launch_bounds__(128)
global_norm_bwd_kernel(const float* __restrict__ a_loc, const float* __restrict__ b_loc, const float* __restrict__ norms,
                       const float* __restrict__ part, int nsplit, int B, int D, float* __restrict__ da,
                       float* __restrict__ db) {
  __shared__ float red[4];
  // launched with the programmatic-dependent-launch attribute: let the next kernel (sparc_bwd3, which needs this kernel's
  // output only in its last pass) start now, then wait for the kernel that wrote `part` to finish
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x, dir = blockIdx.y;
  const float* x = (dir ? b_loc : a_loc) + (size_t)r * D;
  float* dx = (dir ? db : da) + (size_t)r * D;
  const float inv = 1.f / norms[(size_t)dir * B + r];
  float g[8];                                   // D <= 1024 per 128 threads
  float dot = 0.f;
  int cnt = 0;
  for (int d = threadIdx.x; d < D; d += 128, ++cnt) {
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += part[(((size_t)dir * nsplit + k) * B + r) * D + d];
    g[cnt] = s;
    dot = fmaf(s, x[d] * inv, dot);
  }
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  dot = red[0] + red[1] + red[2] + red[3];
  cnt = 0;
  for (int d = threadIdx.x; d < D; d += 128, ++cnt) dx[d] = (g[cnt] - x[d] * inv * dot) * inv;
}

static int gf_splits(int B, int Bg) {
  const int rb = (B + kGR - 1) / kGR, nt = (Bg + kGC - 1) / kGC;
  int s = (296 + 2 * rb - 1) / (2 * rb);
  if (s > nt) s = nt;
  if (s < 1) s = 1;
  return s;
}
static int gb_splits(int B, int Bg, int D) {
  const int rb = (B + kBR - 1) / kBR, nt = (Bg + kBC - 1) / kBC, dz = (D + kBDz - 1) / kBDz;
  int s = (296 + 2 * rb * dz - 1) / (2 * rb * dz);
  if (s > nt) s = nt;
  if (s > 16) s = 16;
  if (s < 1) s = 1;
  return s;
}

// tensor-core implementation (global_infonce_tc.cu)
bool global_tc_supported(int B, int Bg, int D);
size_t global_tc_workspace_bytes(int B, int Bg, int D);
int global_tc_fwd(const float* a_loc, const float* b_loc, const float* a_all, const float* b_all, int B, int Bg, int D,
                  int col_offset, float scale, float eps, float* norms2, float** part_m, float** part_l, float** diag,
                  int* nsplit, void* ws, int gathered_ranks, const PeerTable* peers, cudaStream_t st);
int global_tc_bwd(int B, int Bg, int D, int col_offset, float scale, const float* lse_loc2, const float* lse_all2,
                  const float* coef2, float** dpart, int* nsplit, void* ws, int gathered_ranks, cudaStream_t st);

// low-latency symmetric CUDA-core implementation for the rank-local case (global_infonce_sym.cu)
bool global_sym_supported(int B, int Bg, int D);
size_t global_sym_workspace_bytes(int B, int D);
int global_sym_fwd(const float* a, const float* b, int B, int D, float scale, float eps, float* norms2, float* lse2,
                   float* sums2, const float* local_partial, const uint8_t* mask, int T, float gw, float lw, float* out8,
                   void* ws, cudaStream_t st);
int global_sym_bwd(const float* a, const float* b, int B, int D, float scale, float eps, const float* lse2, const float* coef2,
                   float** dpart, int* nsplit, void* ws, cudaStream_t st);

// path 0 (auto): tensor cores when the shape allows (D % 64 == 0, D <= 512) -- measured under CUDA-graph replay of the
// SPARC step the tensor-core chain beats the symmetric fp32 tiles at every rank-local batch (B = 64: 0.133 vs 0.144 ms per
// step, 256: 0.237 vs 0.248, 512: 0.416 vs 0.491; tools/ab_global.sh) -- else the symmetric fp32 tiles for rank-local
// problems up to B = 512; path 1 keeps everything on fp32 CUDA cores.  CFA_GLOBAL_PREFER_SYM=1 restores the round-1 order.
static bool prefer_sym() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("CFA_GLOBAL_PREFER_SYM"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}
static bool use_sym(int B, int Bg, int D, int path) {
  if (path == 2 || !global_sym_supported(B, Bg, D)) return false;
  return path == 1 || prefer_sym() || !global_tc_supported(B, Bg, D);
}
static bool use_tc(int B, int Bg, int D, int path) { return path != 1 && !use_sym(B, Bg, D, path) && global_tc_supported(B, Bg, D); }

}  // namespace cfa

using namespace cfa;

extern "C" size_t cfa_global_infonce_workspace_bytes(int B, int Bg, int D) {
  const size_t fwd = ((size_t)4 * gf_splits(B, Bg) * B + 2 * (size_t)B) * sizeof(float);
  const size_t bwd = (size_t)2 * gb_splits(B, Bg, D) * B * D * sizeof(float);
  size_t n = fwd > bwd ? fwd : bwd;
  if (global_tc_supported(B, Bg, D)) { const size_t t = global_tc_workspace_bytes(B, Bg, D); if (t > n) n = t; }
  if (global_sym_supported(B, Bg, D)) { const size_t t = global_sym_workspace_bytes(B, D); if (t > n) n = t; }
  return n;
}

// path: 0 = auto (tensor cores when D % 64 == 0 and D <= 512), 1 = fp32-exact CUDA cores, 2 = tensor cores
extern "C" int cfa_global_infonce_path(int B, int Bg, int D, int path) {
  if (path == 2 && !global_tc_supported(B, Bg, D)) return CFA_ERR_UNSUPPORTED;
  return use_sym(B, Bg, D, path) ? 3 : (use_tc(B, Bg, D, path) ? 2 : 1);
}

extern "C" int cfa_global_infonce_fwd(const float* a_loc, const float* b_loc, const float* a_all, const float* b_all, int B,
                                      int Bg, int D, int col_offset, float scale, float eps, float* lse2, float* norms2,
                                      float* sums2, const float* local_partial, const uint8_t* mask, int T, float gw,
                                      float lw, float* out8, void* workspace, size_t workspace_bytes, int path,
                                      int gathered_ranks, void* stream) {
  return cfa::global_infonce_fwd_peers(a_loc, b_loc, a_all, b_all, B, Bg, D, col_offset, scale, eps, lse2, norms2, sums2,
                                       local_partial, mask, T, gw, lw, out8, workspace, workspace_bytes, path, gathered_ranks,
                                       nullptr, stream);
}

// peers != NULL (peer_exchange.cu): the "all" rows are read from every rank's exchange block through the table; a_all /
// b_all are ignored
int cfa::global_infonce_fwd_peers(const float* a_loc, const float* b_loc, const float* a_all, const float* b_all, int B,
                                  int Bg, int D, int col_offset, float scale, float eps, float* lse2, float* norms2,
                                  float* sums2, const float* local_partial, const uint8_t* mask, int T, float gw, float lw,
                                  float* out8, void* workspace, size_t workspace_bytes, int path, int gathered_ranks,
                                  const PeerTable* peers, void* stream) {
  if (B <= 0 || Bg <= 0 || D <= 0 || col_offset < 0 || col_offset + B > Bg) return CFA_ERR_BAD_ARG;
  if (gathered_ranks > 1 && (gathered_ranks * B != Bg || !use_tc(B, Bg, D, path))) return CFA_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < cfa_global_infonce_workspace_bytes(B, Bg, D)) return CFA_ERR_WORKSPACE;
  if (out8 && (Bg != B || !local_partial || !mask)) return CFA_ERR_BAD_ARG;   // fused scalar epilogue: single process only
  if (path == 2 && !global_tc_supported(B, Bg, D)) return CFA_ERR_UNSUPPORTED;
  if (peers && !use_tc(B, Bg, D, path)) return CFA_ERR_UNSUPPORTED;
  if (use_sym(B, Bg, D, path))          // one launch: the last CTA merges the partials and writes the scalar outputs
    return global_sym_fwd(a_loc, b_loc, B, D, scale, eps, norms2, lse2, sums2, local_partial, mask, T, gw, lw, out8, workspace,
                          (cudaStream_t)stream);
  if (use_tc(B, Bg, D, path)) {
    float *pm, *pl, *dg;
    int nsp;
    const int rc = global_tc_fwd(a_loc, b_loc, a_all, b_all, B, Bg, D, col_offset, scale, eps, norms2, &pm, &pl, &dg, &nsp,
                                 workspace, gathered_ranks, peers, (cudaStream_t)stream);
    if (rc != CFA_OK) return rc;
    if (nsp > 8) {                                    // measured at 32 partials per row: single merge CTA 15 us, grid merge + sum 4 + 2 us
      CFA_CUDA_TRY(cfa_launch_pdl(2, global_merge_rows_kernel, dim3((2 * B + 31) / 32), dim3(256), 0, (cudaStream_t)stream,
                                  (const float*)pm, (const float*)pl, B, nsp, lse2));
      nsp = 0;
    }
    CFA_CUDA_TRY(cfa_launch_pdl(2, global_combine_kernel, dim3(1), dim3(kNT), 0, (cudaStream_t)stream, (const float*)pm,
                                (const float*)pl, (const float*)dg, B, nsp, lse2, sums2, Bg, local_partial, mask, T, gw, lw, out8));
    return launch_status();
  }
  const int ns = gf_splits(B, Bg);
  float* part_m = (float*)workspace;
  float* part_l = part_m + (size_t)2 * ns * B;
  float* diag = part_l + (size_t)2 * ns * B;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((B + kGR - 1) / kGR, ns, 2);
  const size_t fsmem = sizeof(float) * ((kGR + kGC) * kGLd + kGR + kGC);
  CFA_SMEM_ATTR_ONCE(global_fwd_kernel, fsmem);
  global_fwd_kernel<<<grid, kNT, fsmem, st>>>(a_loc, b_loc, a_all, b_all, B, Bg, D, col_offset, scale, eps, part_m, part_l,
                                          diag, norms2);
  CFA_CUDA_TRY(cudaGetLastError());
  global_combine_kernel<<<1, kNT, 0, st>>>(part_m, part_l, diag, B, ns, lse2, sums2, Bg, local_partial, mask, T, gw, lw, out8);
  return launch_status();
}

extern "C" int cfa_global_infonce_bwd(const float* a_loc, const float* b_loc, const float* a_all, const float* b_all, int B,
                                      int Bg, int D, int col_offset, float scale, float eps, const float* lse_loc2,
                                      const float* lse_all2, const float* norms2, const float* coef2, float* da, float* db,
                                      void* workspace, size_t workspace_bytes, int path, int gathered_ranks, void* stream) {
  if (B <= 0 || Bg <= 0 || D <= 0 || D > 1024 || col_offset < 0 || col_offset + B > Bg) return CFA_ERR_BAD_ARG;
  if (gathered_ranks > 1 && (gathered_ranks * B != Bg || !use_tc(B, Bg, D, path))) return CFA_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < cfa_global_infonce_workspace_bytes(B, Bg, D)) return CFA_ERR_WORKSPACE;
  if (path == 2 && !global_tc_supported(B, Bg, D)) return CFA_ERR_UNSUPPORTED;
  if (use_sym(B, Bg, D, path)) {
    float* dpart;
    int nsp;
    const int rc = global_sym_bwd(a_loc, b_loc, B, D, scale, eps, lse_loc2, coef2, &dpart, &nsp, workspace, (cudaStream_t)stream);
    if (rc != CFA_OK) return rc;
    CFA_CUDA_TRY(cfa_launch_pdl(1, global_norm_bwd_kernel, dim3(B, 2), dim3(128), 0, (cudaStream_t)stream, a_loc, b_loc, norms2,
                                (const float*)dpart, nsp, B, D, da, db));
    return launch_status();
  }
  if (use_tc(B, Bg, D, path)) {       // needs the SAME workspace the forward call used (normalised hi/lo operands live there)
    float* dpart;
    int nsp;
    const int rc = global_tc_bwd(B, Bg, D, col_offset, scale, lse_loc2, lse_all2, coef2, &dpart, &nsp, workspace,
                                 gathered_ranks, (cudaStream_t)stream);
    if (rc != CFA_OK) return rc;
    CFA_CUDA_TRY(cfa_launch_pdl(1, global_norm_bwd_kernel, dim3(B, 2), dim3(128), 0, (cudaStream_t)stream, a_loc, b_loc, norms2,
                                (const float*)dpart, nsp, B, D, da, db));
    return launch_status();
  }
  const int ns = gb_splits(B, Bg, D);
  const size_t smem = sizeof(float) * (2 * kBR * kBLd + kBR * kBLdS + kBC * kBLdB + kBC);
  CFA_SMEM_ATTR_ONCE(global_bwd_kernel, smem);
  cudaStream_t st = (cudaStream_t)stream;
  const int rb = (B + kBR - 1) / kBR;
  dim3 grid(2 * rb, ns, (D + kBDz - 1) / kBDz);
  global_bwd_kernel<<<grid, kNT, smem, st>>>(a_loc, b_loc, a_all, b_all, B, Bg, D, col_offset, scale, eps, lse_loc2, lse_all2,
                                             norms2, coef2, (float*)workspace);
  CFA_CUDA_TRY(cudaGetLastError());
  global_norm_bwd_kernel<<<dim3(B, 2), 128, 0, st>>>(a_loc, b_loc, norms2, (const float*)workspace, ns, B, D, da, db);
  return launch_status();
}
