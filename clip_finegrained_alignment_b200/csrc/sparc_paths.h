// Internal: the two implementations behind cfa_sparc_fwd / cfa_sparc_bwd.
#pragma once
#include "common.cuh"

namespace cfa {

// fp32-exact CUDA-core path (losses_simt.cu): any dtype, P limited by shared memory
int sparc_fwd_simt(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, int dtype, float thr,
                   float scale, float* pooled_v, float* pooled_l, float* lse_row, float* lse_col, float* local_partial,
                   void* scratch, size_t scratch_bytes, void* stream);
int sparc_bwd_simt(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, int dtype, float thr,
                   float scale, const float* lse_row, const float* lse_col, const float* coef, const float* dpooled_v,
                   const float* dpooled_l, void* dv, void* dl, void* scratch, size_t scratch_bytes, void* stream);

// tcgen05 path (sparc_tc.cu): bf16, D % 256 == 0, P <= 256, T <= 128
bool sparc_tc_supported(int P, int T, int D, int dtype);
int sparc_fwd_tc_launch(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, float thr,
                        float scale, float* row_inv_norm, float* pooled_v, float* pooled_l, float* lse_row,
                        float* lse_col, float* local_partial, float* tt_logits, float* g_inv_norm, void* g_split,
                        float* q_save, cudaStream_t st);

// second-generation tcgen05 forward (sparc_tc_fwd2.cu): prep fused, 8 epilogue warps
bool sparc_fwd2_supported(int P, int T, int D, int dtype);
int sparc_fwd2_launch(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, float thr,
                      float scale, float* row_inv_norm, float* pooled_v, float* pooled_l, float* lse_row, float* lse_col,
                      float* local_partial, float* tt_logits, float* g_inv_norm, void* g_split, float* q_save,
                      long long* prof, int dtype, cudaStream_t st);

// restructured tcgen05 backward (sparc_tc_bwd2.cu): needs the forward's saved G (bf16 hi|lo) and Q = G . v^T
bool sparc_bwd2_supported(int P, int T, int D, int dtype);
int sparc_bwd2_launch(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, float thr, float scale,
                      const float* row_inv_norm, const float* lse_row, const float* lse_col, const float* coef,
                      const float* tt_logits, const float* g_inv_norm, const void* g_split, const float* q_save,
                      const float* dpv, const float* dpl, void* dv, void* dl, long long* prof, int dtype, cudaStream_t st);

// third-generation ("transposed") tcgen05 kernels (sparc_tc_fwd3.cu / sparc_tc_bwd3.cu): raw tiles are the A operand, on-chip
// hi|lo operands are stacked along N; T <= 80, P <= 256, D % 128 == 0.  The forward leaves the per-token statistics
// (min, 1/range, sigma, arg-min patch) in the first B*T*4 floats of the q_save buffer instead of Q.
bool sparc_fwd3_supported(int P, int T, int D, int dtype);
bool sparc_bwd3_supported(int P, int T, int D, int dtype);
bool sparc_gen3_enabled(int P, int T, int D, int dtype);
int sparc_fwd3_launch(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, float thr,
                      float scale, float* row_inv_norm, float* pooled_v, float* pooled_l, float* lse_row, float* lse_col,
                      float* local_partial, float* tt_logits, float* g_inv_norm, void* g_split, float* stats,
                      long long* prof, int dtype, cudaStream_t st);
int sparc_bwd3_launch(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, float thr, float scale,
                      const float* row_inv_norm, const float* lse_row, const float* lse_col, const float* coef,
                      const float* tt_logits, const float* g_inv_norm, const void* g_split, const float* stats,
                      const float* dpv, const float* dpl, void* dv, void* dl, long long* prof, int dtype, cudaStream_t st,
                      bool pdl_late = false);
// Set by the one-call entry points (sparc_fused_abi.cu) around their cfa_sparc_bwd call: the kernel launched just before
// sparc_bwd3 there is global_norm_bwd_kernel, whose only output (d pooled) the backward needs in its last pass, so the
// backward may start under it (programmatic dependent launch) and wait late.  Everywhere else the wait is the kernel's
// first instruction.
extern thread_local bool g_sparc_bwd_pdl_late;
// Non-null while a one-call entry point runs its backward: gt_bwd_kernel and sparc_bwd3_kernel then evaluate the upstream-
// gradient fan-in themselves (CoefSrc, common.cuh) and no coefficient kernel is launched.
extern thread_local const CoefSrc* g_coef_src;

// cfa_global_infonce_fwd with the gathered rows read through a peer table (global_infonce.cu)
int global_infonce_fwd_peers(const float* a_loc, const float* b_loc, const float* a_all, const float* b_all, int B, int Bg,
                             int D, int col_offset, float scale, float eps, float* lse2, float* norms2, float* sums2,
                             const float* local_partial, const uint8_t* mask, int T, float gw, float lw, float* out8,
                             void* workspace, size_t workspace_bytes, int path, int gathered_ranks, const PeerTable* peers,
                             void* stream);

}  // namespace cfa
