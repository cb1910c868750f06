"""Stage-by-stage check of the third-generation tensor-core SPARC kernels (cfa_sparc_fwd / cfa_sparc_bwd through the
C ABI) against fp64 torch math: row norms, pooled means, saved G (hi + lo), statistics, logits, LSE; then the full
SPARCLoss against the oracle.  A hung kernel is reported with the clock64 phase stamps it left in pinned host memory
instead of blocking the process.   usage: python tools/dbg_gen3.py [B P T D [f16]]"""
import os, sys, time
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
import torch
sys.path.insert(0, '.')
from clip_finegrained_alignment_b200 import _lib

B, P, T, D = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (3, 196, 77, 512)
DT = torch.float16 if (len(sys.argv) >= 6 and sys.argv[5] == "f16") else torch.bfloat16
s, thr = 1.0, float(torch.tensor(1.0 / P, dtype=torch.float32))
g = torch.Generator().manual_seed(1)
v = torch.randn(B, P, D, generator=g).to(DT)
l = torch.randn(B, T, D, generator=g).to(DT)
m = torch.ones(B, T, dtype=torch.bool)
dev = torch.device("cuda")
vv, ll, mm = v.to(dev), l.to(dev), m.to(dev).view(torch.uint8)
code = _lib.DTYPE_CODE[DT]
L = _lib.lib
print("path", L.cfa_sparc_path(P, T, D, code, 0), "bwd path", L.cfa_sparc_bwd_path(P, T, D, code, 0))

prof = torch.zeros(B * 32, dtype=torch.int64).pin_memory()
_lib.call("cfa_debug_set_profile_buffer_fwd", prof.data_ptr())

NP = (P + 15) & ~15
sizes = (2 * B * D, 8, B * T, B * T, 2 * B, B * (P + T), B * T * T, B * T, B * T * D, B * T * NP)
off = [0]
for n in sizes:
    off.append(off[-1] + ((n + 31) & ~31))
blk = torch.full((off[-1],), float("nan"), dtype=torch.float32, device=dev)
ptr = [blk.data_ptr() + 4 * o for o in off]
torch.cuda.synchronize()
ev = torch.cuda.Event()
_lib.call("cfa_sparc_fwd", vv.data_ptr(), ll.data_ptr(), mm.data_ptr(), B, P, T, D, code, thr, s,
          ptr[5], ptr[0], ptr[0] + 4 * B * D, ptr[2], ptr[3], ptr[4], ptr[6], ptr[7], ptr[8], ptr[9], 0, 0, 0, _lib.stream_ptr())
ev.record()
t0 = time.time()
while not ev.query() and time.time() - t0 < 8:
    time.sleep(0.05)
if not ev.query():
    print("FWD KERNEL HUNG; stamps (MMA thread [0..15], epilogue thread 0 [16..31]) of sample 0:")
    st = prof[:32].tolist()
    base = st[0]
    print(" mma:", [x - base if x else None for x in st[:8]])
    print(" epi:", [x - base if x else None for x in st[16:28]])
    os._exit(1)
torch.cuda.synchronize()
print("fwd done in %.3f s" % (time.time() - t0))
st = prof[:32].tolist(); base = st[0]
print(" mma stamps:", [x - base for x in st[:4]], " epi stamps:", [x - base for x in st[16:26]])

# ---------------- reference math (fp64)
vd, ld = v.double(), l.double()
vn = vd.norm(dim=-1).clamp_min(1e-12); ln = ld.norm(dim=-1).clamp_min(1e-12)
Sraw = torch.einsum('btd,bpd->btp', ld, vd)
S = Sraw / ln[:, :, None] / vn[:, None, :]
mn = S.min(-1, keepdim=True)[0]; mx = S.max(-1, keepdim=True)[0]
N = (S - mn) / (mx - mn + 1e-8)
Th = torch.where(N < thr, torch.zeros_like(N), N)
sig = Th.sum(-1, keepdim=True).clamp_min(1e-8)
W = Th / sig
G = torch.einsum('btp,bpd->btd', W, vd)
gn = G.norm(dim=-1).clamp_min(1e-12)
logits = s * torch.einsum('btd,bjd->btj', G / gn[..., None], ld / ln[..., None])
lse_r = torch.logsumexp(logits, dim=2); lse_c = torch.logsumexp(logits, dim=1)

def cmp(name, got, ref):
    got = got.double().cpu(); ref = ref.double()
    err = (got - ref).abs().max().item(); rel = ((got - ref).norm() / ref.norm().clamp_min(1e-300)).item()
    print(f"  {name:12s} max abs {err:.3e}  rel {rel:.3e}  finite {bool(torch.isfinite(got).all())}")

rin = blk[off[5]:off[5] + B * (P + T)]
cmp("inv_vn", rin[:B * P].view(B, P), 1 / vn)
cmp("inv_ln", rin[B * P:].view(B, T), 1 / ln)
pooled = blk[:2 * B * D].view(2, B, D)
cmp("pooled_v", pooled[0], vd.mean(1))
cmp("pooled_l", pooled[1], ld.mean(1))
gs = blk[off[8]:off[8] + B * T * D].view(DT).view(B, 2, T, D).float()
cmp("G", gs[:, 0] + gs[:, 1], G)
cmp("g_inv_norm", blk[off[7]:off[7] + B * T].view(B, T), 1 / gn)
stt = blk[off[9]:off[9] + B * T * 4].view(B, T, 4)
cmp("min", stt[..., 0], mn[..., 0]); cmp("1/range", stt[..., 1], 1 / (mx - mn + 1e-8)[..., 0]); cmp("sigma", stt[..., 2], sig[..., 0])
print("  argmin ok:", bool((stt[..., 3].contiguous().view(torch.int32).cpu() == S.argmin(-1).int()).all()))
cmp("logits", blk[off[6]:off[6] + B * T * T].view(B, T, T), logits)
cmp("lse_row", blk[off[2]:off[2] + B * T].view(B, T), lse_r)
cmp("lse_col", blk[off[3]:off[3] + B * T].view(B, T), lse_c)

# ---------------- full loss + backward against the oracle
import types
from clip_finegrained_alignment_b200 import SPARCLoss
from oracle import losses_oracle as lo
prof2 = torch.zeros(B * 32, dtype=torch.int64).pin_memory()
_lib.call("cfa_debug_set_profile_buffer", prof2.data_ptr())
cfg = types.SimpleNamespace(similarity_threshold=thr, global_loss_weight=0.9, local_loss_weight=1.1, inverse_temperature=s)
v2 = vv.clone().requires_grad_(True); l2 = ll.clone().requires_grad_(True)
out = SPARCLoss(cfg, kernel_path="tc")(v2, l2, m.to(dev))
out["total_loss"].backward()
ev2 = torch.cuda.Event(); ev2.record()
t0 = time.time()
while not ev2.query() and time.time() - t0 < 8:
    time.sleep(0.05)
if not ev2.query():
    print("BWD KERNEL HUNG; stamps of sample 0:")
    st = prof2[:32].tolist(); base = st[0]
    print(" mma:", [x - base if x else None for x in st[:8]])
    print(" epi:", [x - base if x else None for x in st[16:28]])
    os._exit(1)
torch.cuda.synchronize()
st = prof2[:32].tolist(); base = st[0]
print(" bwd mma stamps:", [x - base for x in st[:5]], " epi stamps:", [x - base for x in st[16:24]])
o = lo.sparc_forward(v.double(), l.double(), m, thr, 0.9, 1.1, s)
rv, rl = lo.sparc_backward(o)
for k in lo.SPARC_KEYS:
    print(f"  {k:14s} {float(out[k]):.7f} vs {float(o[k]):.7f}")
cmp("dv", v2.grad.float(), rv)
cmp("dl", l2.grad.float(), rl)
