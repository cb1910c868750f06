// AdamSPD multi-tensor step for sm_100a.  Replaces AdamSPD.adam/_ratio (finetune/optimizers.py:100-157).
//
// HBM-bound elementwise work: pass 1 streams p,g,m,v,pre (20 B/elt) and writes p,m,v (12 B/elt) with
// 128-bit accesses; three per-tensor reductions ride along (warp shuffles -> one fp64 atomic per CTA).
// Pass 2 touches only tensors whose device-side condition says "project" (+12 B/elt).  No host sync.
#include "common.cuh"

namespace cfa {

constexpr int kAdamThreads = 256;
constexpr int kAdamVecPerThread = 8;                                      // float4 per thread per chunk
constexpr int kAdamChunk = kAdamThreads * 4 * kAdamVecPerThread;          // 8192 elements

__device__ __forceinline__ float4 ld_stream(const float4* p) {            // read-only, no L1 allocation
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_rw(const float4* p) {                // read-then-overwritten streams
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w) : "memory");
}

struct AdamScalars {
  float beta1, omb1, beta2, omb2, eps, step_size, sqrt_bc2;
};

// One element, reference evaluation order (optimizers.py:128-143).  IEEE sqrt/div (no fast-math).
__device__ __forceinline__ void adam_elem(const AdamScalars& h, float p, float g, float& m, float& v, float* vmax,
                                          float pre, float& new_p, float& s_cond, float& s_new, float& s_old) {
  m = fmaf(g, h.omb1, m * h.beta1);                   // exp_avg.mul_(b1).add_(grad, alpha=1-b1)
  v = fmaf(h.omb2 * g, g, v * h.beta2);               // exp_avg_sq.mul_(b2).addcmul_(grad, grad, value=1-b2)
  float vv = v;
  if (vmax) { *vmax = fmaxf(*vmax, v); vv = *vmax; }  // amsgrad (:131-135)
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv), h.sqrt_bc2), h.eps);
  new_p = __fsub_rn(p, __fdiv_rn(__fmul_rn(h.step_size, m), denom));
  const float d_old = __fsub_rn(p, pre);
  const float d_new = __fsub_rn(new_p, pre);
  s_cond = fmaf(g, d_old, s_cond);                    // sum grad*(param-pre)        (:147)
  s_new = fmaf(d_new, d_new, s_new);                  // ||new_p-pre||^2             (:155)
  s_old = fmaf(d_old, d_old, s_old);                  // ||param-pre||^2             (:155)
}

// kAmp: gradients are consumed as (g * inv_scale) * clip_coef (GradScaler.unscale_ followed by clip_grad_norm_,
// finetuner.py:150-151, same two fp32 roundings) and the whole step is skipped when a non-finite gradient was found
// (what GradScaler.step does on the host, finetuner.py:152) — amp = {inv_scale, clip_coef, total_norm, found_inf}.
template <bool kAms, bool kAmp>
__global__ void __launch_bounds__(kAdamThreads)
adamspd_pass1(const cfa_adamspd_tensor* __restrict__ tensors, const cfa_adamspd_chunk* __restrict__ chunks,
              double* __restrict__ reduce, const float* __restrict__ amp) {
  float gs0 = 1.f, gs1 = 1.f;
  if (kAmp) {
    if (amp[3] != 0.f) return;
    gs0 = amp[0]; gs1 = amp[1];
  }
  const cfa_adamspd_chunk ck = chunks[blockIdx.x];
  const cfa_adamspd_tensor t = tensors[ck.tensor];
  const AdamScalars h{t.beta1, t.one_minus_beta1, t.beta2, t.one_minus_beta2, t.eps, t.step_size, t.sqrt_bc2};
  const int64_t base = (int64_t)ck.chunk * kAdamChunk;
  const int64_t remain = t.numel - base;
  const int n = remain < kAdamChunk ? (int)remain : kAdamChunk;

  float* p = (float*)t.p + base;
  const float* g = (const float*)t.g + base;
  float* m = (float*)t.m + base;
  float* v = (float*)t.v + base;
  const float* pre = t.pre ? (const float*)t.pre + base : nullptr;
  float* vmax = kAms ? (float*)t.vmax + base : nullptr;

  double a_cond = 0.0, a_new = 0.0, a_old = 0.0;
  const bool aligned = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v) |
                         ((uintptr_t)pre) | ((uintptr_t)vmax)) & 15) == 0;
  if (aligned && n == kAdamChunk) {
    // full chunk, 128-bit path: two batches of 4 float4 per stream keep 20-24 loads in flight per thread
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float4 P[4], G[4], M[4], V[4], R[4], X[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = (half * 4 + u) * kAdamThreads + threadIdx.x;
        P[u] = ld_rw((const float4*)p + idx);
        G[u] = ld_stream((const float4*)g + idx);
        if (kAmp) {
          G[u].x = __fmul_rn(__fmul_rn(G[u].x, gs0), gs1); G[u].y = __fmul_rn(__fmul_rn(G[u].y, gs0), gs1);
          G[u].z = __fmul_rn(__fmul_rn(G[u].z, gs0), gs1); G[u].w = __fmul_rn(__fmul_rn(G[u].w, gs0), gs1);
        }
        M[u] = ld_rw((const float4*)m + idx);
        V[u] = ld_rw((const float4*)v + idx);
        R[u] = pre ? ld_stream((const float4*)pre + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (kAms) X[u] = ld_rw((const float4*)vmax + idx);
      }
      float s_cond = 0.f, s_new = 0.f, s_old = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = (half * 4 + u) * kAdamThreads + threadIdx.x;
        float4 np;
        adam_elem(h, P[u].x, G[u].x, M[u].x, V[u].x, kAms ? &X[u].x : nullptr, R[u].x, np.x, s_cond, s_new, s_old);
        adam_elem(h, P[u].y, G[u].y, M[u].y, V[u].y, kAms ? &X[u].y : nullptr, R[u].y, np.y, s_cond, s_new, s_old);
        adam_elem(h, P[u].z, G[u].z, M[u].z, V[u].z, kAms ? &X[u].z : nullptr, R[u].z, np.z, s_cond, s_new, s_old);
        adam_elem(h, P[u].w, G[u].w, M[u].w, V[u].w, kAms ? &X[u].w : nullptr, R[u].w, np.w, s_cond, s_new, s_old);
        st_stream((float4*)p + idx, np);
        st_stream((float4*)m + idx, M[u]);
        st_stream((float4*)v + idx, V[u]);
        if (kAms) st_stream((float4*)vmax + idx, X[u]);
      }
      a_cond += (double)s_cond; a_new += (double)s_new; a_old += (double)s_old;
    }
  } else {
    // ragged tail / unaligned views: scalar path
    for (int i0 = threadIdx.x; i0 < n; i0 += kAdamThreads * 4) {
      float s_cond = 0.f, s_new = 0.f, s_old = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kAdamThreads;
        if (i < n) {
          float mm = m[i], vv = v[i], np;
          float vm = kAms ? vmax[i] : 0.f;
          const float gi = kAmp ? __fmul_rn(__fmul_rn(g[i], gs0), gs1) : g[i];
          adam_elem(h, p[i], gi, mm, vv, kAms ? &vm : nullptr, pre ? pre[i] : 0.f, np, s_cond, s_new, s_old);
          p[i] = np; m[i] = mm; v[i] = vv;
          if (kAms) vmax[i] = vm;
        }
      }
      a_cond += (double)s_cond; a_new += (double)s_new; a_old += (double)s_old;
    }
  }

  // block reduction: warp shuffles, then one fp64 atomic per quantity per CTA
  __shared__ double red[3][kAdamThreads / kWarp];
  a_cond = warp_sum(a_cond); a_new = warp_sum(a_new); a_old = warp_sum(a_old);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { red[0][w] = a_cond; red[1][w] = a_new; red[2][w] = a_old; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kAdamThreads / kWarp; ++i) s += red[threadIdx.x][i];
    atomicAdd(reduce + 3 * (int64_t)ck.tensor + threadIdx.x, s);
  }
}

// ---- AMP prologue: GradScaler.unscale_ + clip_grad_norm_ folded into the step ----------------
// pass 0: per-tensor sum of squares of the UNSCALED gradient g * inv_scale and the non-finite check of
// torch._amp_foreach_non_finite_check_and_unscale_ (on the raw value) — one read of g (4 B/elt), nothing written back.
__global__ void __launch_bounds__(kAdamThreads)
adamspd_gradnorm(const cfa_adamspd_tensor* __restrict__ tensors, const cfa_adamspd_chunk* __restrict__ chunks,
                 const float* __restrict__ grad_scale, double* __restrict__ gsum /* [n_tensors] */, float* __restrict__ amp) {
  const cfa_adamspd_chunk ck = chunks[blockIdx.x];
  const cfa_adamspd_tensor t = tensors[ck.tensor];
  const float inv = grad_scale ? (float)(1.0 / (double)*grad_scale) : 1.f;      // _scale.double().reciprocal().float()
  const int64_t base = (int64_t)ck.chunk * kAdamChunk;
  const int64_t remain = t.numel - base;
  const int n = remain < kAdamChunk ? (int)remain : kAdamChunk;
  const float* g = (const float*)t.g + base;
  float ss = 0.f;
  bool bad = false;
  if ((((uintptr_t)g) & 15) == 0 && n == kAdamChunk) {
    float4 G[kAdamVecPerThread];
#pragma unroll
    for (int u = 0; u < kAdamVecPerThread; ++u) G[u] = __ldg((const float4*)g + u * kAdamThreads + threadIdx.x);
#pragma unroll
    for (int u = 0; u < kAdamVecPerThread; ++u) {
      const float x[4] = {G[u].x, G[u].y, G[u].z, G[u].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bad |= !isfinite(x[j]);
        const float y = __fmul_rn(x[j], inv);
        ss = fmaf(y, y, ss);
      }
    }
  } else {
    for (int i = threadIdx.x; i < n; i += kAdamThreads) {
      const float x = g[i];
      bad |= !isfinite(x);
      const float y = __fmul_rn(x, inv);
      ss = fmaf(y, y, ss);
    }
  }
  __shared__ double red[kAdamThreads / kWarp];
  double a = warp_sum((double)ss);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = a;
  const bool any_bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kAdamThreads / kWarp; ++i) s += red[i];
    atomicAdd(gsum + ck.tensor, s);
    if (any_bad) amp[3] = 1.f;
  }
}

// one CTA: total_norm = || (||g_t||)_t ||_2 as clip_grad_norm_ forms it (per-tensor fp32 norms, then the norm of the stack),
// clip_coef = min(max_norm / (total_norm + 1e-6), 1)   (torch/nn/utils/clip_grad.py)
__global__ void __launch_bounds__(kAdamThreads)
adamspd_clipcoef(const double* __restrict__ gsum, int n_tensors, const float* __restrict__ grad_scale, float max_norm,
                 float* __restrict__ amp) {
  __shared__ double red[kAdamThreads / kWarp];
  double a = 0.0;
  for (int i = threadIdx.x; i < n_tensors; i += kAdamThreads) {
    const float nt = sqrtf((float)gsum[i]);
    a += (double)nt * (double)nt;
  }
  a = warp_sum(a);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kAdamThreads / kWarp; ++i) s += red[i];
    const float total = (float)sqrt(s);
    float coef = 1.f;
    if (max_norm > 0.f) coef = fminf(__fdiv_rn(max_norm, __fadd_rn(total, 1e-6f)), 1.f);
    amp[0] = grad_scale ? (float)(1.0 / (double)*grad_scale) : 1.f;
    amp[1] = coef;
    amp[2] = total;
  }
}

// ratio = hardtanh((||new_p-pre|| - ||p-pre||) / ||new_p-pre||, 0, 1) in fp32, NaN passes through (:154-157)
__device__ __forceinline__ bool spd_decide(const double* r3, float& ratio) {
  const float cond = -(float)r3[0];                       // condition = -sum(grad*(param-pre))  (:147)
  ratio = 0.f;
  if (!(cond < 0.0f)) return false;                       // (:148)
  const float curr = sqrtf((float)r3[1]);
  const float prev = sqrtf((float)r3[2]);
  float r = __fdiv_rn(__fsub_rn(curr, prev), curr);
  if (r == r) r = fminf(fmaxf(r, 0.0f), 1.0f);
  ratio = r;
  return true;
}

__global__ void __launch_bounds__(kAdamThreads)
adamspd_pass2(const cfa_adamspd_tensor* __restrict__ tensors, const cfa_adamspd_chunk* __restrict__ chunks,
              const double* __restrict__ reduce, float* __restrict__ stats) {
  const cfa_adamspd_chunk ck = chunks[blockIdx.x];
  float ratio;
  const bool project = spd_decide(reduce + 3 * (int64_t)ck.tensor, ratio);
  if (stats && ck.chunk == 0 && threadIdx.x == 0) {
    stats[2 * ck.tensor] = project ? 1.f : 0.f;
    stats[2 * ck.tensor + 1] = ratio;
  }
  if (!project) return;
  const cfa_adamspd_tensor t = tensors[ck.tensor];
  const float f = __fmul_rn(t.weight_decay, ratio);       // weight_decay * ratio                (:150)
  if (f == 0.0f) return;                                  // new_p - 0*(...) == new_p: nothing to write
  const int64_t base = (int64_t)ck.chunk * kAdamChunk;
  const int64_t remain = t.numel - base;
  const int n = remain < kAdamChunk ? (int)remain : kAdamChunk;
  float* p = (float*)t.p + base;
  const float* pre = t.pre ? (const float*)t.pre + base : nullptr;
  const bool aligned = ((((uintptr_t)p) | ((uintptr_t)pre)) & 15) == 0;
  if (aligned && n == kAdamChunk) {
    float4 P[kAdamVecPerThread], R[kAdamVecPerThread];
#pragma unroll
    for (int u = 0; u < kAdamVecPerThread; ++u) {
      const int idx = u * kAdamThreads + threadIdx.x;
      P[u] = ld_rw((const float4*)p + idx);
      R[u] = pre ? ld_stream((const float4*)pre + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kAdamVecPerThread; ++u) {
      const int idx = u * kAdamThreads + threadIdx.x;
      float4 o;
      o.x = __fsub_rn(P[u].x, __fmul_rn(f, __fsub_rn(P[u].x, R[u].x)));
      o.y = __fsub_rn(P[u].y, __fmul_rn(f, __fsub_rn(P[u].y, R[u].y)));
      o.z = __fsub_rn(P[u].z, __fmul_rn(f, __fsub_rn(P[u].z, R[u].z)));
      o.w = __fsub_rn(P[u].w, __fmul_rn(f, __fsub_rn(P[u].w, R[u].w)));
      st_stream((float4*)p + idx, o);
    }
  } else {
    for (int i = threadIdx.x; i < n; i += kAdamThreads) {
      const float x = p[i];
      p[i] = __fsub_rn(x, __fmul_rn(f, __fsub_rn(x, pre ? pre[i] : 0.f)));
    }
  }
}

}  // namespace cfa

extern "C" int cfa_adamspd_chunk_elems(void) { return cfa::kAdamChunk; }

extern "C" int cfa_adamspd_step(const cfa_adamspd_tensor* d_tensors, int n_tensors, const cfa_adamspd_chunk* d_chunks,
                                int n_chunks, double* d_reduce, float* d_stats, int dtype, int amsgrad, void* stream) {
  if (dtype != CFA_DTYPE_F32) return CFA_ERR_UNSUPPORTED;
  if (n_tensors < 0 || n_chunks < 0 || (n_tensors > 0 && (!d_tensors || !d_chunks || !d_reduce))) return CFA_ERR_BAD_ARG;
  if (n_tensors == 0 || n_chunks == 0) return CFA_OK;
  const bool ams = amsgrad != 0;
  cudaStream_t st = (cudaStream_t)stream;
  CFA_CUDA_TRY(cudaMemsetAsync(d_reduce, 0, sizeof(double) * 3 * (size_t)n_tensors, st));
  if (ams) cfa::adamspd_pass1<true, false><<<n_chunks, cfa::kAdamThreads, 0, st>>>(d_tensors, d_chunks, d_reduce, nullptr);
  else cfa::adamspd_pass1<false, false><<<n_chunks, cfa::kAdamThreads, 0, st>>>(d_tensors, d_chunks, d_reduce, nullptr);
  CFA_CUDA_TRY(cudaGetLastError());
  cfa::adamspd_pass2<<<n_chunks, cfa::kAdamThreads, 0, st>>>(d_tensors, d_chunks, d_reduce, d_stats);
  return cfa::launch_status();
}

extern "C" int cfa_adamspd_step_amp(const cfa_adamspd_tensor* d_tensors, int n_tensors, const cfa_adamspd_chunk* d_chunks,
                                    int n_chunks, double* d_reduce, float* d_stats, const float* d_grad_scale,
                                    float max_grad_norm, float* d_amp, int dtype, int amsgrad, void* stream) {
  if (dtype != CFA_DTYPE_F32) return CFA_ERR_UNSUPPORTED;
  if (n_tensors < 0 || n_chunks < 0 || !d_amp || (n_tensors > 0 && (!d_tensors || !d_chunks || !d_reduce))) return CFA_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  CFA_CUDA_TRY(cudaMemsetAsync(d_amp, 0, 4 * sizeof(float), st));
  if (n_tensors == 0 || n_chunks == 0) return CFA_OK;
  const bool ams = amsgrad != 0;
  CFA_CUDA_TRY(cudaMemsetAsync(d_reduce, 0, sizeof(double) * 4 * (size_t)n_tensors, st));
  double* gsum = d_reduce + 3 * (size_t)n_tensors;
  cfa::adamspd_gradnorm<<<n_chunks, cfa::kAdamThreads, 0, st>>>(d_tensors, d_chunks, d_grad_scale, gsum, d_amp);
  cfa::adamspd_clipcoef<<<1, cfa::kAdamThreads, 0, st>>>(gsum, n_tensors, d_grad_scale, max_grad_norm, d_amp);
  if (ams) cfa::adamspd_pass1<true, true><<<n_chunks, cfa::kAdamThreads, 0, st>>>(d_tensors, d_chunks, d_reduce, d_amp);
  else cfa::adamspd_pass1<false, true><<<n_chunks, cfa::kAdamThreads, 0, st>>>(d_tensors, d_chunks, d_reduce, d_amp);
  CFA_CUDA_TRY(cudaGetLastError());
  cfa::adamspd_pass2<<<n_chunks, cfa::kAdamThreads, 0, st>>>(d_tensors, d_chunks, d_reduce, d_stats);
  return cfa::launch_status();
}
