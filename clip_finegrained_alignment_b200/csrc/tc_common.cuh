// sm_100a primitives used by the tensor-core kernels: mbarrier, TMA (cp.async.bulk.tensor), tcgen05
// (TMEM alloc, UMMA descriptors, mma, commit, ld) — thin inline-PTX wrappers, no library code.
#pragma once
#include "common.cuh"
#include <cuda.h>   // CUtensorMap (types only; the encode entry point is fetched at run time)
#include <cstdio>
#include <cstring>

namespace cfa {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// same with a suspend-time hint (ns): the waiting thread sleeps in hardware until the phase completes (or the hint runs
// out) instead of re-issuing the probe — spinning warps otherwise take issue slots from the warps that do the work
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
  return ok != 0;
}
// Bounded wait for the third-generation kernels: the probe suspends the thread in hardware; the loop around it is three
// instructions (a spinning warp takes issue slots from the warps that do the work) and the time-out path is an out-of-line
// call: inlined at ~40 wait sites, the printf / trap sequence was a quarter of the kernels' code (instruction fetch).
static __device__ __noinline__ void mbar_timeout_trap(uint32_t bar, uint32_t parity) {
  printf("cfa: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, bar, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t n = 0;
  while (!mbar_try_wait_hint(bar, parity, 100000u)) {
    if (++n > (1u << 26)) mbar_timeout_trap(smem_u32(bar), parity);       // >= 1 s of wake-ups: a protocol bug, not a wait
  }
}

// Bounded wait: a protocol bug traps after ~2 s instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("cfa: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// ---------------------------------------------------------------- proxies / fences
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// ---------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 32 lanes x 16 / 32 consecutive fp32 columns: thread i of the warp reads TMEM lane (lane_base + i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                 "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                 "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
                 "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
                 "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* r) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(r);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the same wait, tied to the registers of a load issued EARLIER than the previous statement (software-pipelined TMEM
// reads: the load of chunk c + 1 is in flight while chunk c is processed): the in/out operands keep the compiler from
// moving any read of r[] above the wait
__device__ __forceinline__ void tmem_ld_wait8(float* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]), "+f"(r[7]) :: "memory");
}
__device__ __forceinline__ void tmem_ld_wait4(float* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]) :: "memory");
}
// registers -> TMEM: thread i of the warp writes 16 consecutive fp32 columns of TMEM lane (lane_base + i)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* r) {
  const uint32_t* u = reinterpret_cast<const uint32_t*>(r);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]),
                 "r"(u[8]), "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// dynamic shared memory base rounded up to 1024 B (SWIZZLE_128B tiles) WITHOUT leaving the shared address space:
// pointer arithmetic on the __shared__ array keeps LDS / STS (a uintptr_t round trip degrades them to generic LD / ST)
#define CFA_SMEM_BASE_1024(arr) ((arr) + ((1024u - (cfa::tc::smem_u32(arr) & 1023u)) & 1023u))

// 8 floats -> bf16 hi (x ~ hi + lo) and lo parts, packed conversions (F2FP.BF16.F32.PACK_AB)
__device__ __forceinline__ void split_hilo8(const float* x, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
    const uint32_t hb = *reinterpret_cast<const uint32_t*>(&hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(x[2 * i] - __uint_as_float(hb << 16),
                                                    x[2 * i + 1] - __uint_as_float(hb & 0xffff0000u));
    h[i] = hb;
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// the same split into fp16 hi | lo parts (fp16 embeddings: tcgen05 kind::f16 takes no mixed fp16 x bf16 operand pair, so
// the on-chip operands follow the raw tiles' format; the callers scale them into fp16's normal range by a power of two)
__device__ __forceinline__ void split_hilo8_f16(const float* x, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
template <bool kHalf>
__device__ __forceinline__ void split_hilo8_t(const float* x, uint4& hi, uint4& lo) {
  if (kHalf) split_hilo8_f16(x, hi, lo);
  else split_hilo8(x, hi, lo);
}
// largest power of two <= x (x > 0, normal), clamped to [2^-60, 2^60]: exact scale factors for the fp16 operand path
__device__ __forceinline__ float pow2_floor_clamped(float x) {
  x = fminf(fmaxf(x, 8.673617379884035e-19f), 1.152921504606847e18f);       // NaN -> lower clamp
  return __int_as_float(__float_as_int(x) & 0x7f800000);
}

// warp-wide min / max of one fp32 value per lane (CREDUX.MIN/MAX.F32, sm_100a): every lane gets the result
__device__ __forceinline__ float warp_redux_min(float v) {
  float r;
  asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float warp_redux_max(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
// Column sums of 8 per-lane values over the 32 lanes of a warp (16 shuffles): lane l ends up with the sum of column
// (l >> 1) & 7 in v[0] (the same value in lanes l, l^1, l^16).
__device__ __forceinline__ float warp_colsum8(float* v, int lane) {
#pragma unroll
  for (int k = 0; k < 8; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], 16);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const bool up = lane & 8;
    const float send = up ? v[k] : v[k + 4], keep = up ? v[k + 4] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const bool up = lane & 4;
    const float send = up ? v[k] : v[k + 2], keep = up ? v[k + 2] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const bool up = lane & 2;
    const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];
}

// Same for 4 per-lane values (9 shuffles): lane l ends up with the sum of column ((l >> 3) & 1) * 2 + ((l >> 2) & 1) in v[0]
// (lanes 0, 4, 8, 12 hold columns 0, 1, 2, 3).
__device__ __forceinline__ float warp_colsum4(float* v, int lane) {
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], 16);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const bool up = lane & 8;
    const float send = up ? v[k] : v[k + 2], keep = up ? v[k + 2] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const bool up = lane & 4;
    const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];
}

// ---------------------------------------------------------------- UMMA descriptors
constexpr uint32_t kLayoutNone = 0, kLayoutSw128 = 2, kLayoutSw64 = 4;

// 64-bit shared-memory matrix descriptor (PTX ISA "tcgen05 shared memory descriptor"):
// [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) swizzle mode.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// 32-bit instruction descriptor for kind::f16, bf16 x bf16 -> fp32, dense.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// same with the two operand formats chosen independently (true = fp16, false = bf16): raw fp16 embedding tiles (what
// torch.autocast hands the loss by default) are multiplied with the bf16 hi/lo operands produced on chip
__host__ __device__ constexpr uint32_t make_idesc16(int M, int N, bool a_mn_major, bool b_mn_major, bool a_f16, bool b_f16) {
  return (1u << 4) | ((a_f16 ? 0u : 1u) << 7) | ((b_f16 ? 0u : 1u) << 10) | ((a_mn_major ? 1u : 0u) << 15) |
         ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// 8 raw 16-bit embedding values (bf16 or fp16) <-> fp32
template <bool kHalf>
__device__ __forceinline__ void unpack_raw8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (kHalf) {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    } else {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
}
template <bool kHalf>
__device__ __forceinline__ uint4 pack_raw8(const float* f) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (kHalf) {
      const __half2 t = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&t);
    } else {
      const __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&t);
    }
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// D[tmem] (+)= A[smem] . B[smem]   (one thread issues)
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// Warp-uniform variants: call from ALL lanes of a converged warp; `leader` (from elect_one(), fixed for the
// whole kernel so that commits track the same thread's MMAs) predicates the instruction itself, so descriptor
// arithmetic stays in the uniform datapath instead of being shuffled into uniform registers per instruction.
__device__ __forceinline__ void umma_ss_w(bool leader, uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate), "r"((uint32_t)leader) : "memory");
}
__device__ __forceinline__ void umma_commit_w(bool leader, uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar)), "r"((uint32_t)leader) : "memory");
}

// arrive on an mbarrier when every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- operand layouts
// SWIZZLE_128B K-major tile of [rows][64] bf16 (what TMA with CU_TENSOR_MAP_SWIZZLE_128B writes): byte offset
// of element (r, k): rows at 128 B, 16-byte chunks XOR-ed with (r % 8).  Tile base 1024-byte aligned.
__device__ __forceinline__ uint32_t sw128_offset(int r, int k) {
  return (uint32_t)(r * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + ((k & 7) << 1));
}
// "Interleaved" (no-swizzle) layout for operands written by threads: [k/8][row][k%8], `rows` a multiple of 8.
//   read K-major  : LBO = rows*16 (next 8-wide k chunk), SBO = 128 (next 8 rows)
//   read MN-major : the same bytes describe the transposed operand: LBO = 128, SBO = rows*16
__device__ __forceinline__ uint32_t il_offset(int rows, int r, int k) {
  return (uint32_t)((k >> 3) * rows * 16 + r * 16 + ((k & 7) << 1));
}

}  // namespace tc

// host: cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency)
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline PFN_tmapEncodeTiled get_tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_tmapEncodeTiled)p;
  }
  return fn;
}

// bf16 tensor [d2][d1][d0] (d0 contiguous), box [1][box1][box0], zero fill out of bounds.
// box0 = 64 elements -> 128-byte swizzle, box0 = 32 elements -> 64-byte swizzle.
// plane_stride (elements, 0 = d0 * d1): distance between consecutive planes when the [d1 x d0] window is a row range of a
// taller matrix (the local rows inside the all-gathered operand array)
static inline int make_tmap_bf16_3d(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box0,
                                    uint32_t box1, bool f16 = false, uint64_t plane_stride = 0) {
  PFN_tmapEncodeTiled enc = get_tmap_encoder();
  if (!enc) return CFA_ERR_UNSUPPORTED;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * 2, (plane_stride ? plane_stride : d0 * d1) * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, box0 == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    // The encoder is a DRIVER entry point: on a thread without a current context (a backward whose first CUDA call is this
    // one, on autograd's worker thread) it reports CUDA_ERROR_INVALID_CONTEXT -> bind the primary context and retry.
    if (r != CUDA_ERROR_INVALID_CONTEXT) break;
    cudaFree(nullptr);
  }
  return r == CUDA_SUCCESS ? CFA_OK : CFA_ERR_BAD_ARG;
}

}  // namespace cfa
