// One C-ABI call per direction for the rank-local SPARC loss (losses.py:199-264): the host side of a training step is
// Python, and at ~0.3 ms of device time per step every ctypes call (~15 us) and every torch allocation (~5 us) counts.
//   cfa_sparc_loss_fwd = cfa_sparc_fwd + cfa_global_infonce_fwd (with the fused scalar epilogue)
//   cfa_sparc_loss_bwd = cfa_sparc_coef_ptrs + cfa_global_infonce_bwd + cfa_sparc_bwd
// over ONE caller-provided workspace whose layout is decided here (the Python side only allocates it and reads the 8
// floats at its start).  The gathered (multi-rank) loss keeps the separate entry points: its two collectives are issued
// by the host between them.
#include "common.cuh"
#include "sparc_paths.h"

namespace cfa {

struct LossWs {
  size_t out8, pooled, lse_row, lse_col, part, rin, tt, gin, gsplit, qsave, glse, gsums, gnorms, coef, dab, gws, scratch, total;
  size_t gpack;            // gathered loss: [world][2B+2] packs pulled from the peers
  size_t gws_bytes, scratch_bytes;
  bool saved;
};

static inline size_t up32(size_t n) { return (n + 31) & ~(size_t)31; }      // 128-byte pieces (floats)

static LossWs loss_ws_layout(int B, int P, int T, int D, int dtype, int path, int world = 1) {
  LossWs w;
  const size_t NP = ((size_t)P + 15) & ~(size_t)15;
  const int which = cfa_sparc_path(P, T, D, dtype, path);
  w.saved = which == 2;                                   // tensor-core path saves G (hi|lo) and Q for the backward
  size_t o = 0;
  w.out8 = o; o += 32;
  w.pooled = o; o += up32((size_t)2 * B * D);
  w.lse_row = o; o += up32((size_t)B * T);
  w.lse_col = o; o += up32((size_t)B * T);
  w.part = o; o += up32((size_t)2 * B);
  w.rin = o; o += up32((size_t)B * (P + T));
  w.tt = o; o += up32((size_t)B * T * T);
  w.gin = o; o += up32((size_t)B * T);
  w.gsplit = o; o += w.saved ? up32((size_t)B * T * D) : 0;
  w.qsave = o; o += w.saved ? up32((size_t)B * T * NP) : 0;
  w.glse = o; o += up32((size_t)2 * B);
  w.gsums = o; o += 32;
  w.gnorms = o; o += up32((size_t)2 * B);
  w.coef = o; o += 32;
  w.dab = o; o += up32((size_t)2 * B * D);
  w.gws_bytes = cfa_global_infonce_workspace_bytes(B, world * B, D);
  w.gws = o; o += up32((w.gws_bytes + 3) / 4);
  w.gpack = o; o += world > 1 ? up32((size_t)world * (2 * B + 2)) : 0;
  w.scratch_bytes = which == 2 ? 0 : cfa_sparc_scratch_bytes(B, P, T, 1);
  w.scratch = o; o += up32((w.scratch_bytes + 3) / 4);
  w.total = o;
  return w;
}

}  // namespace cfa

using namespace cfa;

extern "C" size_t cfa_sparc_loss_workspace_bytes(int B, int P, int T, int D, int dtype, int path) {
  if (B <= 0 || P <= 0 || T <= 0 || D <= 0) return 0;
  return loss_ws_layout(B, P, T, D, dtype, path).total * sizeof(float);
}

extern "C" int cfa_sparc_loss_fwd(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, int dtype,
                                  float thr, float scale, float gw, float lw, void* workspace, size_t workspace_bytes,
                                  int path, void* stream) {
  if (B <= 0 || P <= 0 || T <= 0 || D <= 0 || !v || !l || !mask || !workspace) return CFA_ERR_BAD_ARG;
  if (((uintptr_t)workspace & 127) != 0) return CFA_ERR_BAD_ARG;
  const LossWs w = loss_ws_layout(B, P, T, D, dtype, path);
  if (workspace_bytes < w.total * sizeof(float)) return CFA_ERR_WORKSPACE;
  float* f = (float*)workspace;
  int rc = cfa_sparc_fwd(v, l, mask, B, P, T, D, dtype, thr, scale, f + w.rin, f + w.pooled, f + w.pooled + (size_t)B * D,
                         f + w.lse_row, f + w.lse_col, f + w.part, f + w.tt, f + w.gin, w.saved ? (void*)(f + w.gsplit) : nullptr,
                         w.saved ? f + w.qsave : nullptr, w.scratch_bytes ? (void*)(f + w.scratch) : nullptr, w.scratch_bytes,
                         path, stream);
  if (rc != CFA_OK) return rc;
  // fp32 inputs (and path 1) keep the fp32-exact global kernels; the rank-local problem goes to the symmetric tiles anyway
  const int gpath = (dtype == CFA_DTYPE_F32 || path == 1) ? 1 : 0;
  const float* a = f + w.pooled;
  const float* b = a + (size_t)B * D;
  return cfa_global_infonce_fwd(a, b, a, b, B, B, D, 0, scale, 1e-12f, f + w.glse, f + w.gnorms, f + w.gsums, f + w.part, mask,
                                T, gw, lw, f + w.out8, f + w.gws, w.gws_bytes, gpath, 0, stream);
}

extern "C" int cfa_sparc_loss_bwd(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D, int dtype,
                                  float thr, float scale, float gw, float lw, void* workspace, size_t workspace_bytes,
                                  const float* g_global, const float* g_local, const float* g_total, const float* g_vl,
                                  const float* g_lv, const float* g_vl_local, const float* g_lv_local, void* dv, void* dl,
                                  int path, void* stream) {
  if (B <= 0 || P <= 0 || T <= 0 || D <= 0 || !v || !l || !mask || !workspace || !dv || !dl) return CFA_ERR_BAD_ARG;
  const LossWs w = loss_ws_layout(B, P, T, D, dtype, path);
  if (workspace_bytes < w.total * sizeof(float)) return CFA_ERR_WORKSPACE;
  float* f = (float*)workspace;
  const int gpath = (dtype == CFA_DTYPE_F32 || path == 1) ? 1 : 0;
  // both consumers on tensor cores (gt_bwd_kernel, sparc_bwd3_kernel): they evaluate the upstream-gradient fan-in themselves
  // and no coefficient kernel is launched; otherwise cfa_sparc_coef_ptrs fills the coefficient array as before
  const bool inline_coef = w.saved && cfa_global_infonce_path(B, B, D, gpath) == 2 && cfa_sparc_bwd_path(P, T, D, dtype, path) == 2 &&
                           cfa::sparc_gen3_enabled(P, T, D, dtype);
  const cfa::CoefSrc cs{{g_global, g_local, g_total, g_vl, g_lv, g_vl_local, g_lv_local}, f + w.out8, gw, lw, 1.f, B, 1};
  int rc = CFA_OK;
  if (!inline_coef) {
    rc = cfa_sparc_coef_ptrs(g_global, g_local, g_total, g_vl, g_lv, g_vl_local, g_lv_local, gw, lw, B, f + w.out8, f + w.coef, stream);
    if (rc != CFA_OK) return rc;
  }
  struct CoefScope {                                   // cleared on every exit path
    explicit CoefScope(const cfa::CoefSrc* p) { cfa::g_coef_src = p; }
    ~CoefScope() { cfa::g_coef_src = nullptr; }
  } coef_scope(inline_coef ? &cs : nullptr);
  const float* a = f + w.pooled;
  const float* b = a + (size_t)B * D;
  float* da = f + w.dab;
  float* db = da + (size_t)B * D;
  rc = cfa_global_infonce_bwd(a, b, a, b, B, B, D, 0, scale, 1e-12f, f + w.glse, f + w.glse, f + w.gnorms, f + w.coef, da, db,
                              f + w.gws, w.gws_bytes, gpath, 0, stream);
  if (rc != CFA_OK) return rc;
  cfa::g_sparc_bwd_pdl_late = true;      // the backward may start under the global InfoNCE backward (see sparc_paths.h)
  rc = cfa_sparc_bwd(v, l, mask, B, P, T, D, dtype, thr, scale, f + w.rin, f + w.lse_row, f + w.lse_col, f + w.tt, f + w.gin,
                       w.saved ? (const void*)(f + w.gsplit) : nullptr, w.saved ? f + w.qsave : nullptr, f + w.coef + 2, da, db,
                       dv, dl, w.scratch_bytes ? (void*)(f + w.scratch) : nullptr, w.scratch_bytes, path, stream);
  cfa::g_sparc_bwd_pdl_late = false;
  return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// Gathered (multi-rank) SPARC loss over peer memory: same one-call-per-direction shape as above; the two exchanges of
// the global InfoNCE (pooled embeddings; [lse | CE sums] packs) go through every rank's exchange block (CUDA IPC over
// NVLink, peer_exchange.cu) instead of two host-issued NCCL all-gathers:
//   sparc_fwd -> sync A (push pooled, barrier) -> split kernel reads the peers' rows in place -> logits tiles, merge
//   -> sync B (push pack, barrier, pull all packs) -> scalar epilogue.        The backward has no exchange at all.
// ---------------------------------------------------------------------------------------------------------------
namespace cfa {
int peer_sync(void* const* h_blocks, int world, int rank, uint32_t epoch, const float* push_src, size_t push_off, size_t push_n,
              size_t pull_off, int pull_n, float* pull_dst, cudaStream_t st, float* tail2_sums = nullptr);
size_t peer_slot_words(int B, int D);
constexpr size_t kPeerHeaderWords = 64;
}

extern "C" size_t cfa_peer_exchange_bytes(int B, int D) {
  if (B <= 0 || D <= 0) return 0;
  return (kPeerHeaderWords + 2 * peer_slot_words(B, D)) * sizeof(float);
}

extern "C" size_t cfa_sparc_loss_gathered_workspace_bytes(int B, int P, int T, int D, int dtype, int path, int world) {
  if (B <= 0 || P <= 0 || T <= 0 || D <= 0 || world < 1 || world > kMaxPeers) return 0;
  return loss_ws_layout(B, P, T, D, dtype, path, world).total * sizeof(float);
}

extern "C" int cfa_sparc_loss_gathered_fwd(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D,
                                           int dtype, float thr, float scale, float gw, float lw, void* workspace,
                                           size_t workspace_bytes, int path, int world, int rank,
                                           void* const* h_peer_blocks, uint32_t step, void* stream) {
  if (B <= 0 || P <= 0 || T <= 0 || D <= 0 || !v || !l || !mask || !workspace || !h_peer_blocks) return CFA_ERR_BAD_ARG;
  if (world < 2 || world > kMaxPeers || rank < 0 || rank >= world) return CFA_ERR_BAD_ARG;
  if (((uintptr_t)workspace & 127) != 0) return CFA_ERR_BAD_ARG;
  const int gpath = (dtype == CFA_DTYPE_F32 || path == 1) ? 1 : 0;
  if (cfa_global_infonce_path(B, world * B, D, gpath) != 2) return CFA_ERR_UNSUPPORTED;     // tensor-core logits tiles only
  const LossWs w = loss_ws_layout(B, P, T, D, dtype, path, world);
  if (workspace_bytes < w.total * sizeof(float)) return CFA_ERR_WORKSPACE;
  float* f = (float*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  // From here on BOTH barriers of the step are issued whatever happens locally: a rank that returned early would leave its
  // peers waiting in theirs until the time-out.  The first local error is reported after the second barrier.
  int rc = cfa_sparc_fwd(v, l, mask, B, P, T, D, dtype, thr, scale, f + w.rin, f + w.pooled, f + w.pooled + (size_t)B * D,
                         f + w.lse_row, f + w.lse_col, f + w.part, f + w.tt, f + w.gin, w.saved ? (void*)(f + w.gsplit) : nullptr,
                         w.saved ? f + w.qsave : nullptr, w.scratch_bytes ? (void*)(f + w.scratch) : nullptr, w.scratch_bytes,
                         path, stream);
  const size_t slot = kPeerHeaderWords + (size_t)(step & 1) * peer_slot_words(B, D);
  const size_t pack_off = slot + up32((size_t)2 * B * D);
  // exchange 1: my pooled [2][B][D] rows become readable by every peer
  const int rc1 = peer_sync(h_peer_blocks, world, rank, 2 * step + 1, f + w.pooled, slot, (size_t)2 * B * D, 0, 0, nullptr, st);
  if (rc == CFA_OK) rc = rc1;
  PeerTable pt{};
  pt.n = world;
  for (int r = 0; r < world; ++r) pt.base[r] = (const float*)h_peer_blocks[r] + slot;
  const float* a = f + w.pooled;
  const float* b = a + (size_t)B * D;
  // pack = [lse_a (B) | lse_b (B) | sum CE_a, sum CE_b]: glse and gsums must be adjacent -> use the pack row of this rank
  float* pack = f + w.gpack + (size_t)rank * (2 * B + 2);
  if (rc == CFA_OK)
    rc = global_infonce_fwd_peers(a, b, nullptr, nullptr, B, world * B, D, rank * B, scale, 1e-12f, pack, f + w.gnorms,
                                  pack + 2 * B, nullptr, nullptr, 0, 0.f, 0.f, nullptr, f + w.gws, w.gws_bytes, gpath, world, &pt,
                                  stream);
  // exchange 2: packs of every rank -> gpack (the raw [world][2B+2] layout the finalize / backward kernels index)
  const int rc2 = peer_sync(h_peer_blocks, world, rank, 2 * step + 2, pack, pack_off, (size_t)2 * B + 2, pack_off, 2 * B + 2,
                            f + w.gpack, st);
  if (rc == CFA_OK) rc = rc2;
  if (rc != CFA_OK) return rc;
  return cfa_sparc_finalize(f + w.gpack, world * B, f + w.part, mask, B, T, gw, lw, f + w.out8, world, stream);
}

namespace cfa {
int sparc_coef_ptrs_scaled(const float* const g[7], float gw, float lw, int global_batch, const float* out8, float* coef8,
                           float gscale, cudaStream_t st);
}

extern "C" int cfa_sparc_loss_gathered_bwd(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D,
                                           int dtype, float thr, float scale, float gw, float lw, void* workspace,
                                           size_t workspace_bytes, const float* g_global, const float* g_local,
                                           const float* g_total, const float* g_vl, const float* g_lv,
                                           const float* g_vl_local, const float* g_lv_local, void* dv, void* dl, int path,
                                           int world, int rank, void* stream) {
  return cfa_sparc_loss_gathered_bwd_ex(v, l, mask, B, P, T, D, dtype, thr, scale, gw, lw, workspace, workspace_bytes, g_global,
                                        g_local, g_total, g_vl, g_lv, g_vl_local, g_lv_local, dv, dl, path, world, rank, 1.f, stream);
}

extern "C" int cfa_sparc_loss_gathered_bwd_ex(const void* v, const void* l, const uint8_t* mask, int B, int P, int T, int D,
                                              int dtype, float thr, float scale, float gw, float lw, void* workspace,
                                              size_t workspace_bytes, const float* g_global, const float* g_local,
                                              const float* g_total, const float* g_vl, const float* g_lv,
                                              const float* g_vl_local, const float* g_lv_local, void* dv, void* dl, int path,
                                              int world, int rank, float global_grad_scale, void* stream) {
  if (B <= 0 || P <= 0 || T <= 0 || D <= 0 || !v || !l || !mask || !workspace || !dv || !dl) return CFA_ERR_BAD_ARG;
  if (world < 2 || world > kMaxPeers || rank < 0 || rank >= world) return CFA_ERR_BAD_ARG;
  const LossWs w = loss_ws_layout(B, P, T, D, dtype, path, world);
  if (workspace_bytes < w.total * sizeof(float)) return CFA_ERR_WORKSPACE;
  float* f = (float*)workspace;
  const float* const g7[7] = {g_global, g_local, g_total, g_vl, g_lv, g_vl_local, g_lv_local};
  int rc = sparc_coef_ptrs_scaled(g7, gw, lw, world * B, f + w.out8, f + w.coef, global_grad_scale, (cudaStream_t)stream);
  if (rc != CFA_OK) return rc;
  const int gpath = (dtype == CFA_DTYPE_F32 || path == 1) ? 1 : 0;
  const float* a = f + w.pooled;
  const float* b = a + (size_t)B * D;
  float* da = f + w.dab;
  float* db = da + (size_t)B * D;
  const float* pack = f + w.gpack + (size_t)rank * (2 * B + 2);
  // a_all / b_all are not read by the tensor-core backward (its normalised hi/lo operands were saved by the forward)
  rc = cfa_global_infonce_bwd(a, b, a, b, B, world * B, D, rank * B, scale, 1e-12f, pack, f + w.gpack, f + w.gnorms, f + w.coef,
                              da, db, f + w.gws, w.gws_bytes, gpath, world, stream);
  if (rc != CFA_OK) return rc;
  cfa::g_sparc_bwd_pdl_late = true;      // the backward may start under the global InfoNCE backward (see sparc_paths.h)
  rc = cfa_sparc_bwd(v, l, mask, B, P, T, D, dtype, thr, scale, f + w.rin, f + w.lse_row, f + w.lse_col, f + w.tt, f + w.gin,
                       w.saved ? (const void*)(f + w.gsplit) : nullptr, w.saved ? f + w.qsave : nullptr, f + w.coef + 2, da, db,
                       dv, dl, w.scratch_bytes ? (void*)(f + w.scratch) : nullptr, w.scratch_bytes, path, stream);
  cfa::g_sparc_bwd_pdl_late = false;
  return rc;
}

// ---------------------------------------------------------------------------------------------------------------
// The gathered global InfoNCE on its own over peer memory (CustomCLIPLoss / CLIPCountLoss with gather=True,
// losses.py:14-36 on the all-gathered batch): same exchange and kernels as the global part of the gathered SPARC loss.
//   ab_loc [2][B][D] fp32 raw rows (image | text) -> sums2 [2] = sum over the GLOBAL batch of CE_a, CE_b
// Workspace (floats): norms [2B] | packs [world][2B+2] | global kernel workspace.  The backward needs no exchange.
// ---------------------------------------------------------------------------------------------------------------
namespace cfa {
struct GatherWs { size_t norms, gpack, gws, total, gws_bytes; };
static GatherWs gather_ws_layout(int B, int D, int world) {
  GatherWs w;
  size_t o = 0;
  w.norms = o; o += up32((size_t)2 * B);
  w.gpack = o; o += up32((size_t)world * (2 * B + 2));
  w.gws_bytes = cfa_global_infonce_workspace_bytes(B, world * B, D);
  w.gws = o; o += up32((w.gws_bytes + 3) / 4);
  w.total = o;
  return w;
}
}  // namespace cfa

extern "C" size_t cfa_global_infonce_gathered_workspace_bytes(int B, int D, int world) {
  if (B <= 0 || D <= 0 || world < 1 || world > kMaxPeers) return 0;
  return gather_ws_layout(B, D, world).total * sizeof(float);
}

extern "C" int cfa_global_infonce_gathered_fwd(const float* ab_loc, int B, int D, float scale, float eps, void* workspace,
                                               size_t workspace_bytes, int world, int rank, void* const* h_peer_blocks,
                                               uint32_t step, float* sums2, void* stream) {
  if (B <= 0 || D <= 0 || !ab_loc || !workspace || !h_peer_blocks || !sums2) return CFA_ERR_BAD_ARG;
  if (world < 2 || world > kMaxPeers || rank < 0 || rank >= world || ((uintptr_t)workspace & 127) != 0) return CFA_ERR_BAD_ARG;
  if (cfa_global_infonce_path(B, world * B, D, 0) != 2) return CFA_ERR_UNSUPPORTED;
  const GatherWs w = gather_ws_layout(B, D, world);
  if (workspace_bytes < w.total * sizeof(float)) return CFA_ERR_WORKSPACE;
  float* f = (float*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t slot = kPeerHeaderWords + (size_t)(step & 1) * peer_slot_words(B, D);
  const size_t pack_off = slot + up32((size_t)2 * B * D);
  int rc = peer_sync(h_peer_blocks, world, rank, 2 * step + 1, ab_loc, slot, (size_t)2 * B * D, 0, 0, nullptr, st);
  if (rc != CFA_OK) return rc;
  PeerTable pt{};
  pt.n = world;
  for (int r = 0; r < world; ++r) pt.base[r] = (const float*)h_peer_blocks[r] + slot;
  float* pack = f + w.gpack + (size_t)rank * (2 * B + 2);
  rc = global_infonce_fwd_peers(ab_loc, ab_loc + (size_t)B * D, nullptr, nullptr, B, world * B, D, rank * B, scale, eps, pack,
                                f + w.norms, pack + 2 * B, nullptr, nullptr, 0, 0.f, 0.f, nullptr, f + w.gws, w.gws_bytes, 0, world,
                                &pt, stream);
  // the second barrier is issued even if the launch above failed: the peers are waiting for it
  const int rcx = peer_sync(h_peer_blocks, world, rank, 2 * step + 2, pack, pack_off, (size_t)2 * B + 2, pack_off, 2 * B + 2,
                            f + w.gpack, st, sums2);             // the same kernel adds up the ranks' CE sums
  return rc != CFA_OK ? rc : rcx;
}

extern "C" int cfa_global_infonce_gathered_bwd(const float* ab_loc, int B, int D, float scale, float eps, void* workspace,
                                               size_t workspace_bytes, const float* coef2, float* dab, int world, int rank,
                                               void* stream) {
  if (B <= 0 || D <= 0 || !ab_loc || !workspace || !coef2 || !dab) return CFA_ERR_BAD_ARG;
  if (world < 2 || world > kMaxPeers || rank < 0 || rank >= world) return CFA_ERR_BAD_ARG;
  const GatherWs w = gather_ws_layout(B, D, world);
  if (workspace_bytes < w.total * sizeof(float)) return CFA_ERR_WORKSPACE;
  float* f = (float*)workspace;
  const float* a = ab_loc;
  const float* b = ab_loc + (size_t)B * D;
  const float* pack = f + w.gpack + (size_t)rank * (2 * B + 2);
  return cfa_global_infonce_bwd(a, b, a, b, B, world * B, D, rank * B, scale, eps, pack, f + w.gpack, f + w.norms, coef2, dab,
                                dab + (size_t)B * D, f + w.gws, w.gws_bytes, 0, world, stream);
}
