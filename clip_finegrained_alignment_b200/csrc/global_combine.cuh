// Merge of the per-split online-softmax partials of the global InfoNCE (+ optional SPARC scalar epilogue), as a device
// function: it is the body of global_combine_kernel (global_infonce.cu) and also runs inside the LAST CTA of
// global_sym_fwd_kernel (global_infonce_sym.cu), which saves a dependent single-CTA launch per step.  The partials were
// written by other CTAs (possibly of the same grid): they are read with L2-coherent loads (__ldcg), never through the
// read-only path.
#pragma once
#include "common.cuh"
#include "simt_tile.cuh"
#include <math_constants.h>

namespace cfa {

__device__ __forceinline__ float block_sum(float x, float* red) {
  x = warp_sum(x);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
  __syncthreads();
  float s = 0.f;
  for (int k = 0; k < kNT / 32; ++k) s += red[k];
  return s;
}

// One CTA: merge the per-split online-softmax partials -> lse[2][B], sums2 = (sum CE_a, sum CE_b) over the LOCAL rows.
// If out8 != NULL (single process) it also performs the SPARC scalar epilogue (losses.py:163,196,217,252-264).
static __device__ void global_combine_body(const float* part_m, const float* part_l,
                                    const float* diag, int B, int nsplit, float* lse,
                                    float* __restrict__ sums2, int global_batch, const float* __restrict__ local_partial,
                                    const uint8_t* __restrict__ mask, int T, float gw, float lw, float* __restrict__ out8,
                                    float* red /* shared [8] */) {
  float ce[2] = {0.f, 0.f};
  for (int idx = threadIdx.x; idx < 2 * B; idx += kNT) {
    const int dir = idx / B, i = idx - dir * B;
    if (nsplit == 0) {                                // rows already merged by global_merge_rows_kernel: lse holds the result
      ce[dir] += __ldcg(lse + idx) - __ldcg(diag + idx);
      continue;
    }
    float M = -CUDART_INF_F;
    for (int s = 0; s < nsplit; ++s) M = fmaxf(M, __ldcg(part_m + ((size_t)dir * nsplit + s) * B + i));
    float Lsum = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const float m = __ldcg(part_m + ((size_t)dir * nsplit + s) * B + i);
      if (m != -CUDART_INF_F) Lsum += __ldcg(part_l + ((size_t)dir * nsplit + s) * B + i) * expf(m - M);
    }
    const float x = M + logf(Lsum);
    lse[idx] = x;
    ce[dir] += x - __ldcg(diag + idx);
  }
  const float sa = block_sum(ce[0], red), sb = block_sum(ce[1], red);
  if (threadIdx.x == 0) { sums2[0] = sa; sums2[1] = sb; }
  if (out8) {
    float nv = 0.f, a = 0.f, c = 0.f;
    const int nb = (int)((size_t)global_batch);     // single process: global_batch == B
    {                                                 // count of valid tokens: 16 mask bytes per load when the layout allows
      const int n = nb * T;
      int cnt = 0;
      if ((((uintptr_t)mask) & 15) == 0) {
        const int n16 = n >> 4;
        for (int i = threadIdx.x; i < n16; i += kNT) {
          const uint4 w = __ldg(reinterpret_cast<const uint4*>(mask) + i);
          cnt += __popc(__vcmpne4(w.x, 0u) & 0x01010101u) + __popc(__vcmpne4(w.y, 0u) & 0x01010101u) +
                 __popc(__vcmpne4(w.z, 0u) & 0x01010101u) + __popc(__vcmpne4(w.w, 0u) & 0x01010101u);
        }
        for (int i = (n16 << 4) + threadIdx.x; i < n; i += kNT) cnt += mask[i] ? 1 : 0;
      } else {
        for (int i = threadIdx.x; i < n; i += kNT) cnt += mask[i] ? 1 : 0;
      }
      nv = (float)cnt;                                // exact: counts < 2^24 per thread
    }
    for (int i = threadIdx.x; i < nb; i += kNT) { a += local_partial[2 * i]; c += local_partial[2 * i + 1]; }
    nv = block_sum(nv, red); a = block_sum(a, red); c = block_sum(c, red);
    if (threadIdx.x == 0) {
      const float n_valid = nv + 1e-8f;               // evaluated in fp32 like the reference (losses.py:196)
      const float vl = sa / (float)global_batch, lv = sb / (float)global_batch;
      const float vll = a / n_valid, lvl = c / n_valid;
      const float g = 0.5f * (vl + lv), lo = 0.5f * (vll + lvl);
      out8[0] = g; out8[1] = lo; out8[2] = gw * g + lw * lo;
      out8[3] = vl; out8[4] = lv; out8[5] = vll; out8[6] = lvl; out8[7] = n_valid;
    }
  }
}


}  // namespace cfa
