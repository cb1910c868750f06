// Internal: the two implementations behind cfa_sparc_fwd / cfa_sparc_bwd.
#pragma once
#include "common.cuh"

namespace cfa {

// fp32-exact CUDA-core path (losses_simt.cu): any dtype, P limited by shared memory
int sparc_fwd_simt(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, int dtype, float thr,
                   float scale, float* pooled_v, float* pooled_l, float* lse_row, float* lse_col, float* local_partial,
                   void* stream);
int sparc_bwd_simt(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, int dtype, float thr,
                   float scale, const float* lse_row, const float* lse_col, const float* coef, const float* dpooled_v,
                   const float* dpooled_l, void* dv, void* dl, void* stream);

// tcgen05 path (sparc_tc.cu): bf16, D % 256 == 0, P <= 256, T <= 128
bool sparc_tc_supported(int P, int T, int D, int dtype);
int sparc_fwd_tc_launch(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, float thr,
                        float scale, float* row_inv_norm, float* pooled_v, float* pooled_l, float* lse_row,
                        float* lse_col, float* local_partial, float* tt_logits, float* g_inv_norm, cudaStream_t st);

}  // namespace cfa
