/*
 * cfa_b200.h — C ABI of libcfa_b200.so: the sm_100a implementation of the contrastive-loss
 * and optimizer hot path of tpeat/clip-finegrained-alignment.
 *
 * The reference has no FFI/plugin layer: its boundary for this path is three Python classes
 * (finetune/losses.py: SPARCLoss :136-264, CustomCLIPLoss :7-36; finetune/optimizers.py: AdamSPD :8-157)
 * whose arithmetic is torch tensor ops.  The entry points below are what a ctypes binding inside those
 * classes calls instead of the torch ops; each one names the reference lines it replaces.
 * INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name starts with h_ (host);
 *   - tensors are dense row-major; sizes are element counts; `dtype` is one of CFA_DTYPE_*;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - no allocation, no host synchronisation and no host read-back inside any call; scratch is
 *     passed in and sized by the matching *_workspace_bytes query;
 *   - return value: 0 = success, >0 = a cudaError_t, <0 = CFA_ERR_*; cfa_error_string() decodes both.
 *   - sm_100a only.  There is no CPU fallback: without a B200 every compute call fails.
 */
#ifndef CFA_B200_H
#define CFA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFA_ABI_VERSION 4

#define CFA_DTYPE_F32 0
#define CFA_DTYPE_BF16 1
#define CFA_DTYPE_F16 2

#define CFA_OK 0
#define CFA_ERR_BAD_ARG (-1)
#define CFA_ERR_UNSUPPORTED (-2)
#define CFA_ERR_WORKSPACE (-3)

int cfa_abi_version(void);
const char* cfa_error_string(int code);

/* ------------------------------------------------------------------------------------------------
 * AdamSPD multi-tensor step — replaces the per-tensor Python loop AdamSPD.adam + _ratio
 * (finetune/optimizers.py:100-157; state handling of :31-98 stays in Python).
 * ---------------------------------------------------------------------------------------------- */
typedef struct cfa_adamspd_tensor {
  void* p;            /* parameter, updated in place                     (optimizers.py:151) */
  const void* g;      /* gradient                                        (optimizers.py:116) */
  void* m;            /* exp_avg, updated in place                       (optimizers.py:128) */
  void* v;            /* exp_avg_sq, updated in place                    (optimizers.py:129) */
  const void* pre;    /* anchor group['pre'][j]; NULL = zeros            (optimizers.py:146) */
  void* vmax;         /* max_exp_avg_sq when amsgrad, else NULL          (optimizers.py:131-135) */
  int64_t numel;
  float beta1, one_minus_beta1;
  float beta2, one_minus_beta2;
  float eps;
  float step_size;    /* lr / (1 - beta1^t), evaluated in double on the host (optimizers.py:123,139) */
  float sqrt_bc2;     /* sqrt(1 - beta2^t), evaluated in double on the host  (optimizers.py:124,137) */
  float weight_decay;
} cfa_adamspd_tensor;   /* 88 bytes, 8-byte aligned */

typedef struct cfa_adamspd_chunk {
  int32_t tensor;     /* index into the tensor table */
  int32_t chunk;      /* chunk number inside that tensor; elements [chunk*chunk_elems, ...) */
} cfa_adamspd_chunk;

/* elements per chunk the kernels expect in the chunk table */
int cfa_adamspd_chunk_elems(void);

/*
 * One optimizer step over every tensor in the table (two launches + one memset, no host sync):
 *   pass 1: moments, bias-corrected update, p <- new_p, and per-tensor sums
 *           sum g*(p-pre), sum (new_p-pre)^2, sum (p-pre)^2            (optimizers.py:128-147,155)
 *   pass 2: for tensors with condition < 0 and ratio > 0: p <- p - wd*ratio*(p - pre)  (:148-150,154-157)
 * d_reduce: [3*n_tensors] doubles of scratch (zeroed by the call).
 * d_stats : optional [2*n_tensors] floats: (projected ? 1 : 0, ratio) per tensor, or NULL.
 * amsgrad : non-zero = every tensor carries vmax (optimizers.py:131-135).
 */
int cfa_adamspd_step(const cfa_adamspd_tensor* d_tensors, int n_tensors,
                     const cfa_adamspd_chunk* d_chunks, int n_chunks,
                     double* d_reduce, float* d_stats, int dtype, int amsgrad, void* stream);

/*
 * The same step with GradScaler.unscale_ and clip_grad_norm_ folded in (the reference issues them as two extra
 * read-modify-write passes over every gradient right before the step, finetune/finetuner.py:150-152):
 *   pass 0 reads g once: per-tensor ||g / scale||^2 and the non-finite check; one CTA forms total_norm and
 *   clip_coef = min(max_grad_norm / (total_norm + 1e-6), 1); pass 1 consumes (g * inv_scale) * clip_coef (the
 *   reference's two fp32 roundings) and is skipped entirely, like GradScaler.step, when a gradient is inf / NaN.
 * Gradients are left untouched in memory.  d_grad_scale: DEVICE scalar (GradScaler's scale) or NULL = 1;
 * max_grad_norm <= 0: no clipping.  d_amp: DEVICE [4] out = { inv_scale, clip_coef, total_norm, found_inf (0/1) }.
 * d_reduce: [4*n_tensors] doubles here (the 4th block holds the gradient sums of squares).
 */
int cfa_adamspd_step_amp(const cfa_adamspd_tensor* d_tensors, int n_tensors, const cfa_adamspd_chunk* d_chunks,
                         int n_chunks, double* d_reduce, float* d_stats, const float* d_grad_scale,
                         float max_grad_norm, float* d_amp, int dtype, int amsgrad, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Global (batch-level) InfoNCE, both directions per call (losses.py:14-36 CustomCLIPLoss; :145-163 and :207-217
 * SPARCLoss.pairwise_contrastive_loss on the pooled embeddings).
 *   a_loc, b_loc [B,D]  : this rank's RAW (un-normalised) fp32 rows (image / text)
 *   a_all, b_all [Bg,D] : the all-gathered rows of every rank (== a_loc, b_loc when Bg == B)
 *   row i of this rank is global row col_offset + i; logits = scale * normalize(rows) . normalize(cols)^T with
 *   F.normalize's eps (1e-12 for SPARC, 0 for CustomCLIPLoss).  The B x Bg logits are never written to memory.
 * Forward outputs: lse2 [2][B] (log-sum-exp of each local row, direction 0 = a rows vs b columns, 1 = b vs a),
 *   norms2 [2][B] (max(|row|, eps)), sums2 [2] = sum over LOCAL rows of CE.  When out8 != NULL (single process,
 *   Bg == B) the SPARC scalar epilogue cfa_sparc_finalize is fused in.
 * Backward: gradient w.r.t. the raw local rows of  coef2[0] * sum_i CE_a(i) + coef2[1] * sum_j CE_b(j)  taken over
 *   the GLOBAL batch (both directions share the logits, so cross-rank terms only need the other direction's
 *   gathered lse vector lse_all2 [2][Bg]).  coef2 is a DEVICE pointer (already divided by the global batch).
 * ---------------------------------------------------------------------------------------------- */
size_t cfa_global_infonce_workspace_bytes(int B, int Bg, int D);
int cfa_global_infonce_fwd(const float* a_loc, const float* b_loc, const float* a_all, const float* b_all, int B, int Bg,
                           int D, int col_offset, float scale, float eps, float* lse2, float* norms2, float* sums2,
                           const float* local_partial, const uint8_t* mask, int T, float gw, float lw, float* out8,
                           void* workspace, size_t workspace_bytes, int path, int gathered_ranks, void* stream);
int cfa_global_infonce_bwd(const float* a_loc, const float* b_loc, const float* a_all, const float* b_all, int B, int Bg,
                           int D, int col_offset, float scale, float eps, const float* lse_loc2, const float* lse_all2,
                           const float* norms2, const float* coef2, float* da, float* db, void* workspace,
                           size_t workspace_bytes, int path, int gathered_ranks, void* stream);
/* gathered_ranks > 1 (tensor-core path only): the "all" arrays are the RAW outputs of the two NCCL all-gathers, so no
 * re-layout kernel runs between the collective and the loss:
 *   forward : a_all / b_all point at rank 0's image / text block inside the gathered [ranks][2][B][D] buffer
 *             (global row r*B + i of a_all lives at a_all + r*2*B*D + i*D);
 *   backward: lse_all2 is the gathered [ranks][2B+2] buffer of per-rank packs [lse_a (B) | lse_b (B) | sum CE_a, sum CE_b].
 * gathered_ranks <= 1: plain [Bg][D] / [2][Bg] arrays. */
/* path: 0 = auto (rank-local problems, Bg == B <= 512: low-latency symmetric fp32 tiles, one logits tile serving both
 * directions; otherwise tcgen05 logits tiles with bf16 hi/lo-split normalised operands when D % 64 == 0 and D <= 512),
 * 1 = fp32-exact CUDA-core tiles, 2 = tensor cores or CFA_ERR_UNSUPPORTED.  cfa_global_infonce_path reports what a
 * request resolves to: 1 = CUDA-core tiles, 2 = tensor cores, 3 = symmetric rank-local tiles.  The backward must be given the SAME
 * workspace (and path) as the forward: the tensor-core path keeps its normalised operands there. */
int cfa_global_infonce_path(int B, int Bg, int D, int path);

/* ------------------------------------------------------------------------------------------------
 * SPARC fine-grained path, one CTA per sample (losses.py:207-212 pooling and :221-252 local loss).
 * v [B,P,D], l [B,T,D] in `dtype`; mask [B,T] bytes (torch.bool).  Masked tokens are skipped
 * ("truncate" semantics; identical to the reference for all-True masks).
 *   pooled_v [B,D] = mean_p v ; pooled_l [B,D] = sum_t m l / max(sum_t m, 1e-8)          (:207-212)
 *   lse_row/lse_col [B,T]: row / column log-sum-exp of the masked T x T logits           (:180-193)
 *   local_partial [B,2]: per-sample sum_t m_t CE for the two directions                  (:196)
 * The T x P similarity, its min-max normalisation, threshold, renormalised weights and the grouped
 * patch embeddings (:225-245) live only in shared memory / registers.
 * ---------------------------------------------------------------------------------------------- */
int cfa_sparc_fwd(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, int dtype,
                  float thr, float scale, float* row_inv_norm, float* pooled_v, float* pooled_l, float* lse_row,
                  float* lse_col, float* local_partial, float* tt_logits, float* g_inv_norm, void* g_split,
                  float* q_save, void* scratch, size_t scratch_bytes, int path, void* stream);

/*
 * Backward of the above.  coef: DEVICE pointer to 2 floats = upstream coefficient of loss_vl_local and
 * loss_lv_local, already divided by n_valid.  dpooled_v / dpooled_l [B,D]: gradient w.r.t. the pooled
 * means (from the global InfoNCE), may be NULL.  dv [B,P,D], dl [B,T,D] are written in `dtype`.
 *
 * Saved between forward and backward (written by cfa_sparc_fwd, read by cfa_sparc_bwd; the CUDA-core path
 * ignores them): row_inv_norm [B*(P+T)] = 1/max(|v_p|,eps) then 1/max(|l_t|,eps); tt_logits [B*T*T] = the masked,
 * scaled token x token logits (fp32, 24 KB per sample — NOT the T x P similarity, which never leaves the SM);
 * g_inv_norm [B*T] = 1/max(|G_t|,eps); g_split [B][2][T][D] 16-bit = the grouped embeddings G as hi | lo planes in the
 * embeddings' format (bf16, or fp16 for CFA_DTYPE_F16; 16-byte aligned, read back through TMA); q_save: a 16-byte
 * aligned fp32 buffer of B*T*NP floats, NP = (P+15)&~15, whose CONTENT is private to the kernel generation that ran the
 * forward (third generation, sparc_tc_fwd3.cu: the per-token statistics [B][T][4] = min, 1/range, sigma, arg-min patch
 * in its first B*T*4 floats; second generation, sparc_tc_fwd2.cu: Q = G . v^T as [B][T][NP]) -- hand the backward the
 * buffers the forward wrote, in the same process and with the same CFA_SPARC_GEN setting.
 * g_split / q_save are optional (both NULL or both set): with them the tensor-core backward runs as pure
 * TMA -> tcgen05 streams (sparc_tc_bwd3.cu / sparc_tc_bwd2.cu), without them the first-generation kernels recompute G per
 * D block (sparc_tc.cu; bf16, D % 256 == 0 only -- shapes only the third generation takes return CFA_ERR_WORKSPACE).
 * path: 0 = auto, 1 = fp32-exact CUDA-core kernels, 2 = tcgen05 tensor-core kernels or CFA_ERR_UNSUPPORTED.  Tensor-core
 * shapes: third generation bf16 / fp16, P <= 256, T <= 80, D % 128 == 0 and a shared-memory layout under 227 KB (e.g.
 * P = 196 / 197, T = 77 at D = 512 ... 1024); older generations bf16, D % 256 == 0, P <= 256, T <= 128.  cfa_sparc_path /
 * cfa_sparc_bwd_path report what `auto` resolves to for a shape.  fp16 embeddings use range-scaled fp16 hi | lo on-chip operands (tcgen05 kind::f16 takes no mixed
 * fp16 x bf16 operand pair); everything else (fp32, P > 256, ...) runs on the CUDA-core kernels.
 */
int cfa_sparc_bwd(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, int dtype,
                  float thr, float scale, const float* row_inv_norm, const float* lse_row, const float* lse_col,
                  const float* tt_logits, const float* g_inv_norm, const void* g_split, const float* q_save,
                  const float* coef, const float* dpooled_v, const float* dpooled_l, void* dv, void* dl, void* scratch,
                  size_t scratch_bytes, int path, void* stream);
int cfa_sparc_path(int P, int T, int D, int dtype, int path);
int cfa_sparc_bwd_path(int P, int T, int D, int dtype, int path);   /* backward: tensor cores need P <= ~224 at T = 77 */

/* Global scratch of the CUDA-core path: when the T x P tiles (S, and dW in the backward) do not fit in shared memory
 * (ViT-L/14@336: P = 576 / 577) they live in a per-CTA slice of `scratch` (stays L2-resident); 0 = not needed. */
size_t cfa_sparc_scratch_bytes(int B, int P, int T, int backward);

/* largest P the SPARC kernels accept for a given T (staging tiles in shared memory) */
int cfa_sparc_max_patches(int T, int backward);

/*
 * Rank-local SPARC loss in ONE call per direction (what SPARCLoss.forward / .backward issue when the global InfoNCE is
 * not all-gathered): cfa_sparc_fwd + cfa_global_infonce_fwd (fused scalar epilogue), and cfa_sparc_coef_ptrs +
 * cfa_global_infonce_bwd + cfa_sparc_bwd, over one 128-byte-aligned workspace of cfa_sparc_loss_workspace_bytes() whose
 * layout is private to the library.  After the forward the first 8 floats of the workspace hold out[0..7] (see
 * cfa_sparc_finalize); the backward must be given the same, untouched workspace.  g_*: DEVICE scalars, the upstream
 * gradients of the 7 outputs (NULL = unused); they must stay valid until the backward's kernels have run (on the
 * tensor-core chain the kernels read them directly and no coefficient kernel is launched).  Saves ~4 host calls and ~4
 * allocations per step.  The backward's kernels are chained with programmatic dependent launch (the fine-grained
 * backward starts under the global InfoNCE backward and waits for it right before its output pass); to anything
 * enqueued after the call the stream behaves as usual.  CFA_PDL=0 disables the overlap.
 */
size_t cfa_sparc_loss_workspace_bytes(int B, int P, int T, int D, int dtype, int path);
int cfa_sparc_loss_fwd(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, int dtype, float thr,
                       float scale, float gw, float lw, void* workspace, size_t workspace_bytes, int path, void* stream);
int cfa_sparc_loss_bwd(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, int dtype, float thr,
                       float scale, float gw, float lw, void* workspace, size_t workspace_bytes, const float* g_global,
                       const float* g_local, const float* g_total, const float* g_vl, const float* g_lv,
                       const float* g_vl_local, const float* g_lv_local, void* dv, void* dl, int path, void* stream);

/*
 * All-gathered SPARC loss over PEER MEMORY (the reference's dist_finetuner.py:164-176 keeps the loss
 * rank-local — the gathered global InfoNCE is this library's extension).  One rank = one process = one GPU of one
 * NVLink/NVSwitch box.  Every rank owns an exchange block (cfa_peer_alloc: cudaMalloc + CUDA IPC handle) that its peers
 * map with cfa_peer_open; h_peer_blocks is a HOST array of `world` device pointers, entry r = rank r's block as mapped
 * in THIS process (entry `rank` = the own allocation).  The fine-grained loss stays rank-local; the global InfoNCE
 * scores the local rows against the rows of every rank, read in place from the peers' HBM, and the per-rank
 * [lse | CE sums] packs travel the same way.  Two in-stream device barriers per forward, none in the backward, no NCCL
 * call and no host synchronisation; a rank that never arrives turns the losses into NaN after 20 s instead of hanging.
 * `step` = number of gathered forwards issued so far on this exchange (identical on every rank).
 * Shapes the tensor-core global InfoNCE does not take (D % 64 != 0, D > 512, fp32 inputs) return CFA_ERR_UNSUPPORTED:
 * the caller then uses cfa_global_infonce_fwd/_bwd with NCCL-gathered buffers.
 */
size_t cfa_peer_exchange_bytes(int B, int D);
int cfa_peer_alloc(size_t bytes, void** dev_ptr, unsigned char handle_out[64]);
int cfa_peer_open(const unsigned char handle[64], void** dev_ptr);
int cfa_peer_close(void* dev_ptr);
int cfa_peer_free(void* dev_ptr);
/* one barrier on its own: push `push_words` floats into the own block at word offset push_off_words, signal `epoch`
 * to every peer, wait for theirs, then pull `pull_words` floats from word offset pull_off_words of EVERY block into
 * pull_dst [world][pull_words].  Word offsets >= 64 (the header holds the flags). */
/* Barriers of this rank that ran into their time-out so far (a peer never arrived within CFA_PEER_TIMEOUT_MS, default
 * 600 000 ms): the losses of such a step are NaN.  Asynchronous device-to-host copy of the counter on `stream` into
 * h_count (pinned memory recommended); the caller decides when to look at it (PeerExchange.check()). */
int cfa_peer_status(const void* own_block, unsigned int* h_count, void* stream);
int cfa_peer_sync(void* const* h_peer_blocks, int world, int rank, uint32_t epoch, const float* push_src,
                  size_t push_off_words, size_t push_words, size_t pull_off_words, int pull_words, float* pull_dst,
                  void* stream);
size_t cfa_sparc_loss_gathered_workspace_bytes(int B, int P, int T, int D, int dtype, int path, int world);
int cfa_sparc_loss_gathered_fwd(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, int dtype,
                                float thr, float scale, float gw, float lw, void* workspace, size_t workspace_bytes,
                                int path, int world, int rank, void* const* h_peer_blocks, uint32_t step, void* stream);
int cfa_sparc_loss_gathered_bwd(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, int dtype,
                                float thr, float scale, float gw, float lw, void* workspace, size_t workspace_bytes,
                                const float* g_global, const float* g_local, const float* g_total, const float* g_vl,
                                const float* g_lv, const float* g_vl_local, const float* g_lv_local, void* dv, void* dl,
                                int path, int world, int rank, void* stream);
/* Same with a factor on the GLOBAL term of the gradient.  cfa_sparc_loss_gathered_bwd returns d(global mean loss)/d(local
 * rows) once per rank.  Under DistributedDataParallel the parameter gradients are AVERAGED over ranks afterwards, which
 * would divide the global term (identical on every rank) by the world size relative to all-gather-with-grad semantics,
 * where the reduce-scatter SUMS the ranks' contributions before the DDP mean.  global_grad_scale = world restores
 * d(global mean loss)/d(theta) after that mean (the rank-local fine-grained term needs no factor: its own 1/N_valid is
 * per rank).  global_grad_scale = 1: the plain gradient w.r.t. the local rows. */
int cfa_sparc_loss_gathered_bwd_ex(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, int dtype,
                                   float thr, float scale, float gw, float lw, void* workspace, size_t workspace_bytes,
                                   const float* g_global, const float* g_local, const float* g_total, const float* g_vl,
                                   const float* g_lv, const float* g_vl_local, const float* g_lv_local, void* dv, void* dl,
                                   int path, int world, int rank, float global_grad_scale, void* stream);

/*
 * The gathered global InfoNCE alone over the same exchange blocks (CustomCLIPLoss / CLIPCountLoss with gather = True:
 * losses.py:14-36 on the all-gathered batch).  ab_loc [2][B][D] fp32 = this rank's raw image | text rows;
 * sums2 [2] (DEVICE) = sum over the GLOBAL batch of the two directions' cross-entropies.  Backward: gradient of
 * coef2[0] * sum CE_a + coef2[1] * sum CE_b w.r.t. the local rows -> dab [2][B][D]; no exchange.  Same workspace for both.
 */
size_t cfa_global_infonce_gathered_workspace_bytes(int B, int D, int world);
int cfa_global_infonce_gathered_fwd(const float* ab_loc, int B, int D, float scale, float eps, void* workspace,
                                    size_t workspace_bytes, int world, int rank, void* const* h_peer_blocks,
                                    uint32_t step, float* sums2, void* stream);
int cfa_global_infonce_gathered_bwd(const float* ab_loc, int B, int D, float scale, float eps, void* workspace,
                                    size_t workspace_bytes, const float* coef2, float* dab, int world, int rank,
                                    void* stream);

/*
 * SPARCLoss.masked_pairwise_contrastive_loss on its own (losses.py:165-197): a, b [B,T,D] in `dtype`, mask [B,T] bytes.
 * One direction: rows of a against the columns of b of the same sample, target = same token index.
 *   out2[0] = sum_b sum_{valid i} CE_i / n_valid,  out2[1] = n_valid = sum(mask) + 1e-8 (fp32);
 *   lse_row [B,T] and out2 are what the backward needs; partial [B] is scratch.
 * Backward: grad = DEVICE scalar d(loss); da, db [B,T,D] are written in `dtype`.  fp32 CUDA-core tiles (the fused
 * tensor-core version of the same computation lives inside cfa_sparc_fwd / cfa_sparc_bwd).
 */
int cfa_masked_pairwise_fwd(const void* a, const void* b, const uint8_t* mask, int B, int T, int D, int dtype, float scale,
                            float* lse_row, float* partial, float* out2, void* stream);
int cfa_masked_pairwise_bwd(const void* a, const void* b, const uint8_t* mask, int B, int T, int D, int dtype, float scale,
                            const float* lse_row, const float* out2, const float* grad, void* da, void* db, void* stream);

/*
 * Counting losses (fp32 CUDA-core kernels, any of the three dtypes in and out).
 * cfa_count_contrastive_*: CountLoss's counterfactual term (losses.py:281-301).  ei, ek [B,D], ek_cf [B,C,D];
 *   rows are L2-normalised (no eps); per_sample[b] = log sum_c exp(e_i.e_cf[c] / T) - e_i.e_k / T; *out = mean_b.
 *   include_pos != 0 also puts exp(e_i.e_k / T) in the denominator (the grouping of CLIPCountLoss.count_loss, :69-86).
 *   Backward: grad = DEVICE scalar d(out); dei, dek [B,D], dek_cf [B,C,D] in `dtype`.
 * cfa_logits_ce_*: CountLoss's CLIP term on caller-provided logits (losses.py:276-279): logits_a, logits_b [B,B];
 *   *out = (mean_i CE(a_i, i) + mean_i CE(b_i, i)) / 2; lse2, ce2: [2][B] floats (lse2 feeds the backward).
 */
int cfa_count_contrastive_fwd(const void* ei, const void* ek, const void* ek_cf, int B, int C, int D, int dtype,
                              float temperature, int include_pos, float* per_sample, float* out, void* stream);
int cfa_count_contrastive_bwd(const void* ei, const void* ek, const void* ek_cf, int B, int C, int D, int dtype,
                              float temperature, int include_pos, const float* grad, void* dei, void* dek, void* dek_cf,
                              void* stream);
int cfa_logits_ce_fwd(const void* logits_a, const void* logits_b, int B, int dtype, float* lse2, float* ce2, float* out,
                      void* stream);
int cfa_logits_ce_bwd(const void* logits_a, const void* logits_b, int B, int dtype, const float* lse2, const float* grad,
                      void* dlogits_a, void* dlogits_b, void* stream);

/*
 * Scalar epilogue (losses.py:163,196,217,252-264): sums the per-row / per-sample partials and writes
 * out[0..6] = global_loss, local_loss, total_loss, loss_vl, loss_lv, loss_vl_local, loss_lv_local and
 * out[7] = n_valid (sum of mask).  global_sums: DEVICE [2] = sum_i CE_vl(i), sum_j CE_lv(j) over the GLOBAL
 * batch (after the cross-rank all-reduce of cfa_global_infonce_fwd's sums2 when distributed).
 */
int cfa_sparc_finalize(const float* global_sums, int global_batch, const float* local_partial,
                       const uint8_t* mask, int B, int T, float gw, float lw, float* out8, int gathered_ranks, void* stream);
/* gathered_ranks > 1: global_sums is the gathered [ranks][2B+2] pack buffer above; the two CE sums are added over ranks. */

/*
 * Upstream gradient of the 7 outputs -> kernel coefficients (device side, no sync).
 * grad7: DEVICE [7] floats, d(loss)/d(out[k]) in the out[] order above (the autograd gradient of the out vector).
 * coef8 = { c_vl/Bg, c_lv/Bg, c_vl_local/n_valid, c_lv_local/n_valid, c_lv/Bg, c_vl/Bg, 0, 0 } where
 *   c_vl = g[loss_vl] + (g[global] + gw*g[total])/2, c_vl_local = g[loss_vl_local] + (g[local] + lw*g[total])/2, ...
 * (coef8+0 feeds the image-direction InfoNCE backward, coef8+4 the text-direction one, coef8+2 cfa_sparc_bwd).
 */
int cfa_sparc_coef(const float* grad7, float gw, float lw, int global_batch, const float* out8, float* coef8,
                   void* stream);
/* same, the upstream gradients given as 7 separate DEVICE scalars in the out[] order (NULL = that output is unused):
 * what torch.autograd hands a Function with 7 outputs — no zero-fill / concatenation launches on the way in */
int cfa_sparc_coef_ptrs(const float* g_global, const float* g_local, const float* g_total, const float* g_vl,
                        const float* g_lv, const float* g_vl_local, const float* g_lv_local, float gw, float lw,
                        int global_batch, const float* out8, float* coef8, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Validation hook (not a reference replacement): one CTA computes D[128 x N] = A . B on the tcgen05 tensor
 * cores with the operand flavours the tensor-core kernels use.  a_mode: 0 = A[128,K] via TMA (SWIZZLE_128B,
 * K-major), 1 = A[128,K] thread-written interleaved K-major, 2 = A^T[K,128] interleaved read MN-major.
 * b_mode: 0 = B[N,K] via TMA K-major, 1 = B^T[K,64] via TMA read MN-major (N = 64), 2 = B[N,K] interleaved
 * K-major, 3 = B^T[K,N] interleaved MN-major.  D[n] = sum_k A[m,k] * B[n,k].  bf16 in, fp32 out.
 * ---------------------------------------------------------------------------------------------- */
/* tuning aid: device buffer [B][32] int64 receiving clock64 phase stamps of the tensor-core backward (NULL = off) */
int cfa_debug_set_profile_buffer(void* device_buffer);
int cfa_debug_set_profile_buffer_fwd(void* device_buffer);   /* same for the tensor-core forward */
/* debugging aid: host-mapped (pinned) int32 buffer [cta][16 warps] receiving progress markers of the tensor-core
 * global InfoNCE backward, readable from the host while a kernel is stuck (NULL = off) */
int cfa_debug_set_marker_buffer(void* mapped_buffer);
int cfa_tc_selftest(int a_mode, int b_mode, int N, int K, const void* A, const void* B, float* D, void* stream);
int cfa_tc_selftest_timed(int a_mode, int b_mode, int N, int K, const void* A, const void* B, float* D, int repeat,
                          long long* d_cycles, void* stream);   /* tcgen05 issue/throughput microbenchmark */

#ifdef __cplusplus
}
#endif
#endif /* CFA_B200_H */
