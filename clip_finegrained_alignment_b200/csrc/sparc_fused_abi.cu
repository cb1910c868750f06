// One C-ABI call per direction for the rank-local SPARC loss (losses.py:199-264): the host side of a training step is
// Python, and at ~0.3 ms of device time per step every ctypes call (~15 us) and every torch allocation (~5 us) counts.
//   cfa_sparc_loss_fwd = cfa_sparc_fwd + cfa_global_infonce_fwd (with the fused scalar epilogue)
//   cfa_sparc_loss_bwd = cfa_sparc_coef_ptrs + cfa_global_infonce_bwd + cfa_sparc_bwd
// over ONE caller-provided workspace whose layout is decided here (the Python side only allocates it and reads the 8
// floats at its start).  The gathered (multi-rank) loss keeps the separate entry points: its two collectives are issued
// by the host between them.
#include "common.cuh"

namespace cfa {

struct LossWs {
  size_t out8, pooled, lse_row, lse_col, part, rin, tt, gin, gsplit, qsave, glse, gsums, gnorms, coef, dab, gws, scratch, total;
  size_t gws_bytes, scratch_bytes;
  bool saved;
};

static inline size_t up32(size_t n) { return (n + 31) & ~(size_t)31; }      // 128-byte pieces (floats)

static LossWs loss_ws_layout(int B, int P, int T, int D, int dtype, int path) {
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
  w.gws_bytes = cfa_global_infonce_workspace_bytes(B, B, D);
  w.gws = o; o += up32((w.gws_bytes + 3) / 4);
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
  int rc = cfa_sparc_coef_ptrs(g_global, g_local, g_total, g_vl, g_lv, g_vl_local, g_lv_local, gw, lw, B, f + w.out8, f + w.coef,
                               stream);
  if (rc != CFA_OK) return rc;
  const int gpath = (dtype == CFA_DTYPE_F32 || path == 1) ? 1 : 0;
  const float* a = f + w.pooled;
  const float* b = a + (size_t)B * D;
  float* da = f + w.dab;
  float* db = da + (size_t)B * D;
  rc = cfa_global_infonce_bwd(a, b, a, b, B, B, D, 0, scale, 1e-12f, f + w.glse, f + w.glse, f + w.gnorms, f + w.coef, da, db,
                              f + w.gws, w.gws_bytes, gpath, 0, stream);
  if (rc != CFA_OK) return rc;
  return cfa_sparc_bwd(v, l, mask, B, P, T, D, dtype, thr, scale, f + w.rin, f + w.lse_row, f + w.lse_col, f + w.tt, f + w.gin,
                       w.saved ? (const void*)(f + w.gsplit) : nullptr, w.saved ? f + w.qsave : nullptr, f + w.coef + 2, da, db,
                       dv, dl, w.scratch_bytes ? (void*)(f + w.scratch) : nullptr, w.scratch_bytes, path, stream);
}
