// tcgen05 / TMA self-test: one CTA computes D[128 x N] = A . B with every operand flavour the SPARC
// tensor-core kernels use (TMA-written SWIZZLE_128B tiles read K-major or MN-major, thread-written
// interleaved tiles read K-major or MN-major).  tests/test_gpu_tc.py checks each mode against torch.
#include "tc_common.cuh"

namespace cfa {
using namespace tc;
typedef __nv_bfloat16 bf16;

__global__ void __launch_bounds__(128, 1)
tc_selftest_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const bf16* __restrict__ A, const bf16* __restrict__ B, float* __restrict__ D, int a_mode,
                   int b_mode, int N, int K, int repeat, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* As = base;                 // 64 KB
  uint8_t* Bs = base + 65536;         // 96 KB
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  // ---- stage operands
  uint32_t tx_bytes = 0;
  if (a_mode == 0) tx_bytes += (K / 64) * 16384;
  if (a_mode == 4) tx_bytes += (K / 32) * 8192;
  if (a_mode == 3) tx_bytes += 2 * K * 128;
  if (b_mode == 0) tx_bytes += (K / 64) * N * 128;
  if (b_mode == 1) tx_bytes += K * 128;
  if (b_mode == 4) tx_bytes += (K / 32) * N * 64;
  if (b_mode == 5) tx_bytes += K * 64;
  if (tid == 0 && tx_bytes) {
    mbar_expect_tx(&bar_tma, tx_bytes);
    if (a_mode == 0) for (int kb = 0; kb < K / 64; ++kb) tma_load_3d(As + kb * 16384, &tmA, &bar_tma, kb * 64, 0, 0);
    if (a_mode == 3) { tma_load_3d(As, &tmA, &bar_tma, 0, 0, 0); tma_load_3d(As + K * 128, &tmA, &bar_tma, 64, 0, 0); }
    if (b_mode == 0) for (int kb = 0; kb < K / 64; ++kb) tma_load_3d(Bs + kb * N * 128, &tmB, &bar_tma, kb * 64, 0, 0);
    if (b_mode == 1) tma_load_3d(Bs, &tmB, &bar_tma, 0, 0, 0);
    if (a_mode == 4) for (int kb = 0; kb < K / 32; ++kb) tma_load_3d(As + kb * 8192, &tmA, &bar_tma, kb * 32, 0, 0);
    if (b_mode == 4) for (int kb = 0; kb < K / 32; ++kb) tma_load_3d(Bs + kb * N * 64, &tmB, &bar_tma, kb * 32, 0, 0);
    if (b_mode == 5) tma_load_3d(Bs, &tmB, &bar_tma, 0, 0, 0);
  }
  if (a_mode == 1) {          // A [128 x K] -> interleaved, rows = 128
    for (int i = tid; i < 128 * K; i += 128) { const int r = i / K, k = i % K; *(bf16*)(As + il_offset(128, r, k)) = A[i]; }
  } else if (a_mode == 2) {   // At [K x 128] -> interleaved with "rows" = K (k plays the row role)
    for (int i = tid; i < K * 128; i += 128) { const int k = i / 128, m = i % 128; *(bf16*)(As + il_offset(K, k, m)) = A[i]; }
  }
  if (b_mode == 2) {          // B [N x K] -> interleaved, rows = N
    for (int i = tid; i < N * K; i += 128) { const int r = i / K, k = i % K; *(bf16*)(Bs + il_offset(N, r, k)) = B[i]; }
  } else if (b_mode == 3) {   // Bt [K x N] -> interleaved with "rows" = K
    for (int i = tid; i < K * N; i += 128) { const int k = i / N, n = i % N; *(bf16*)(Bs + il_offset(K, k, n)) = B[i]; }
  }
  fence_proxy_async();        // generic-proxy smem writes -> visible to the tensor core (async proxy)
  __syncthreads();
  if (tx_bytes) mbar_wait(&bar_tma, 0);
  tc_fence_after();

  // ---- issue: warp 0, warp-uniform control flow, descriptors advanced by adds on the start-address field
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, N, a_mode == 2 || a_mode == 3, b_mode == 1 || b_mode == 3 || b_mode == 5);
    // per k-step increments (units of 16 B) and per-K-block jumps for the TMA-tiled flavours
    uint64_t da0, db0;
    uint32_t a_ks, b_ks, a_blk = 0, b_blk = 0, a_per = 1 << 30, b_per = 1 << 30;
    if (a_mode == 0) { da0 = make_smem_desc(smem_u32(As), 16, 1024, kLayoutSw128); a_ks = 2; a_per = 4; a_blk = 16384 >> 4; }
    else if (a_mode == 4) { da0 = make_smem_desc(smem_u32(As), 16, 512, kLayoutSw64); a_ks = 2; a_per = 2; a_blk = 8192 >> 4; }
    else if (a_mode == 1) { da0 = make_smem_desc(smem_u32(As), 128 * 16, 128, kLayoutNone); a_ks = (2 * 128 * 16) >> 4; }
    // At [K x 128] as two TMA sub-tiles [K x 64] (SWIZZLE_128B) read MN-major: LBO = distance between the two 64-wide atoms
    else if (a_mode == 3) { da0 = make_smem_desc(smem_u32(As), (uint32_t)K * 128, 1024, kLayoutSw128); a_ks = 2048 >> 4; }
    else { da0 = make_smem_desc(smem_u32(As), 128, K * 16, kLayoutNone); a_ks = 256 >> 4; }
    if (b_mode == 0) { db0 = make_smem_desc(smem_u32(Bs), 16, 1024, kLayoutSw128); b_ks = 2; b_per = 4; b_blk = (N * 128) >> 4; }
    else if (b_mode == 4) { db0 = make_smem_desc(smem_u32(Bs), 16, 512, kLayoutSw64); b_ks = 2; b_per = 2; b_blk = (N * 64) >> 4; }
    else if (b_mode == 1) { db0 = make_smem_desc(smem_u32(Bs), 16, 1024, kLayoutSw128); b_ks = 2048 >> 4; }
    else if (b_mode == 5) { db0 = make_smem_desc(smem_u32(Bs), 16, 512, kLayoutSw64); b_ks = 1024 >> 4; }
    else if (b_mode == 2) { db0 = make_smem_desc(smem_u32(Bs), N * 16, 128, kLayoutNone); b_ks = (2 * N * 16) >> 4; }
    else { db0 = make_smem_desc(smem_u32(Bs), 128, K * 16, kLayoutNone); b_ks = 256 >> 4; }
    const int nks = K / 16;
    const long long t0 = clock64();
    if (repeat > 1) {                       // throughput probe: same operands every time, no address arithmetic
#pragma unroll 8
      for (int i = 0; i < repeat * nks; ++i)       // cycle through 4 consecutive k-steps: distinct operand addresses, cheap arithmetic
        umma_ss_w(leader, tmem, da0 + (uint32_t)(i & 3) * a_ks, db0 + (uint32_t)(i & 3) * b_ks, idesc, true);
    } else
    for (int rep = 0; rep < repeat; ++rep)
      for (int ks = 0; ks < nks; ++ks) {
        const uint64_t da = da0 + (ks / a_per) * a_blk + (ks % a_per) * a_ks;
        const uint64_t db = db0 + (ks / b_per) * b_blk + (ks % b_per) * b_ks;
        umma_ss_w(leader, tmem, da, db, idesc, ks > 0);
      }
    umma_commit_w(leader, &bar_mma);
    mbar_wait(&bar_mma, 0);
    if (cycles && leader) *cycles = clock64() - t0;
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();

  // ---- read back: warp w owns TMEM lanes 32w .. 32w+31
  const int row = tid;
  for (int c0 = 0; c0 < N; c0 += 16) {
    float r[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[row * N + c0 + j] = r[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace cfa

using namespace cfa;

// Debug / validation entry point (not part of the product surface; declared in include/cfa_b200.h).
extern "C" int cfa_tc_selftest_timed(int a_mode, int b_mode, int N, int K, const void* A, const void* B, float* D,
                                     int repeat, long long* d_cycles, void* stream);

extern "C" int cfa_tc_selftest(int a_mode, int b_mode, int N, int K, const void* A, const void* B, float* D, void* stream) {
  return cfa_tc_selftest_timed(a_mode, b_mode, N, K, A, B, D, 1, nullptr, stream);
}

// same, issuing the whole K loop `repeat` times back to back and reporting the issuing thread's clock64 span
extern "C" int cfa_tc_selftest_timed(int a_mode, int b_mode, int N, int K, const void* A, const void* B, float* D,
                                     int repeat, long long* d_cycles, void* stream) {
  if (N % 16 || N < 16 || N > 256 || K % 16 || K < 16 || K > 256) return CFA_ERR_BAD_ARG;
  if ((a_mode == 0 || b_mode == 0) && K % 64) return CFA_ERR_BAD_ARG;
  if ((a_mode == 4 || b_mode == 4) && K % 32) return CFA_ERR_BAD_ARG;
  if (b_mode == 1 && N != 64) return CFA_ERR_BAD_ARG;
  if (b_mode == 5 && N != 32) return CFA_ERR_BAD_ARG;
  if (a_mode == 3 && K % 8) return CFA_ERR_BAD_ARG;
  if ((b_mode == 2 || b_mode == 3) && (size_t)N * K * 2 > 98304) return CFA_ERR_BAD_ARG;
  CUtensorMap tmA, tmB;
  memset(&tmA, 0, sizeof(tmA));
  memset(&tmB, 0, sizeof(tmB));
  int rc;
  if (a_mode == 0 && (rc = make_tmap_bf16_3d(&tmA, A, K, 128, 1, 64, 128)) != CFA_OK) return rc;
  if (a_mode == 3 && (rc = make_tmap_bf16_3d(&tmA, A, 128, K, 1, 64, K)) != CFA_OK) return rc;
  if (b_mode == 0 && (rc = make_tmap_bf16_3d(&tmB, B, K, N, 1, 64, N)) != CFA_OK) return rc;
  if (b_mode == 1 && (rc = make_tmap_bf16_3d(&tmB, B, 64, K, 1, 64, K)) != CFA_OK) return rc;
  if (a_mode == 4 && (rc = make_tmap_bf16_3d(&tmA, A, K, 128, 1, 32, 128)) != CFA_OK) return rc;
  if (b_mode == 4 && (rc = make_tmap_bf16_3d(&tmB, B, K, N, 1, 32, N)) != CFA_OK) return rc;
  if (b_mode == 5 && (rc = make_tmap_bf16_3d(&tmB, B, 32, K, 1, 32, K)) != CFA_OK) return rc;
  const size_t smem = 65536 + 98304 + 1024;
  CFA_CUDA_TRY(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tmA, tmB, (const bf16*)A, (const bf16*)B, D, a_mode, b_mode, N, K, repeat, d_cycles);
  return launch_status();
}
