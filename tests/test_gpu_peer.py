"""All-gathered SPARC loss over peer memory (csrc/peer_exchange.cu, cfa_sparc_loss_gathered_fwd/_bwd).

One GPU is enough for parity: N "ranks" live in one process, each with its own exchange block and its own CUDA
stream, so the device barriers, the in-place reads of the other ranks' pooled rows and the pack exchange run exactly
as across processes (the block table simply holds N local allocations instead of IPC mappings).  Oracle = the
single-process loss on the concatenated batch for the global part + the per-rank fine-grained loss.
The real two-process run over CUDA IPC (needs >= 2 GPUs) is the last test.
"""
import ctypes as C
import os
import subprocess
import sys
import types

import pytest
import torch

from conftest import ROOT, rel_err
from oracle import losses_oracle as lo

pytestmark = pytest.mark.gpu


class _LocalExchange:
    """N exchange blocks in one process (what PeerExchange builds across processes with CUDA IPC)."""

    def __init__(self, N, B, D):
        from clip_finegrained_alignment_b200 import _lib
        self.lib = _lib
        self.N = N
        nbytes = _lib.lib.cfa_peer_exchange_bytes(B, D)
        assert nbytes >= 4 * (64 + 2 * (2 * B * D + 2 * B + 2))
        self.blocks = (C.c_void_p * N)()
        for r in range(N):
            p, h = C.c_void_p(), (C.c_ubyte * 64)()
            _lib.check(_lib.lib.cfa_peer_alloc(nbytes, C.byref(p), h), "cfa_peer_alloc")
            assert any(h), "IPC handle must be filled in"
            self.blocks[r] = p.value

    def close(self):
        for r in range(self.N):
            self.lib.lib.cfa_peer_free(self.blocks[r])


def test_peer_sync_push_barrier_pull():
    """The barrier kernel on its own: 3 'ranks' on 3 streams push a vector, meet, and pull everybody's."""
    from clip_finegrained_alignment_b200 import _lib
    N, n = 3, 1000
    ex = _LocalExchange(N, 64, 64)
    try:
        streams = [torch.cuda.Stream() for _ in range(N)]
        src = [torch.full((n,), float(r + 1), device="cuda") + torch.arange(n, device="cuda") * 1e-3 for r in range(N)]
        dst = [torch.zeros(N, n, device="cuda") for _ in range(N)]
        torch.cuda.synchronize()
        for epoch in (1, 2, 3):                      # repeated use of the same flags: monotonic epochs
            for r in range(N):
                with torch.cuda.stream(streams[r]):
                    src[r] += 1.0
                    _lib.call("cfa_peer_sync", ex.blocks, N, r, epoch, src[r].data_ptr(), 64, n, 64, n, dst[r].data_ptr(),
                              _lib.stream_ptr())
            torch.cuda.synchronize()
            want = torch.stack(src)
            for r in range(N):
                assert torch.equal(dst[r], want), (epoch, r)
    finally:
        ex.close()


def _load_kernels(B, P, T, D, N, dtype=torch.bfloat16):
    """CUDA loads kernels lazily, and loading one synchronises the context: with all 'ranks' in ONE process that would
    wait for a barrier kernel that is itself waiting for the rank whose launch is being loaded.  (Across processes the
    peers keep running, so there is no such dependency.)  Run every kernel of the gathered path once, rank-locally."""
    from clip_finegrained_alignment_b200 import SPARCLoss, _lib
    L = _lib.lib
    v = torch.randn(B, P, D, device="cuda").to(dtype).requires_grad_(True)
    l = torch.randn(B, T, D, device="cuda").to(dtype).requires_grad_(True)
    m = torch.ones(B, T, dtype=torch.bool, device="cuda")
    cfg = types.SimpleNamespace(similarity_threshold=1.0 / P, global_loss_weight=1.0, local_loss_weight=1.0,
                                inverse_temperature=1.0)
    SPARCLoss(cfg)(v, l, m)["total_loss"].backward()
    Bg = N * B
    a = torch.randn(Bg, D, device="cuda"); b = torch.randn(Bg, D, device="cuda")
    nb = L.cfa_global_infonce_workspace_bytes(B, Bg, D)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    pack = torch.zeros(N, 2 * B + 2, device="cuda"); n2 = torch.empty(2, B, device="cuda")
    coef = torch.zeros(8, device="cuda"); dab = torch.empty(2, B, D, device="cuda"); out8 = torch.empty(8, device="cuda")
    part = torch.zeros(B, 2, device="cuda")
    _lib.call("cfa_global_infonce_fwd", a.data_ptr(), b.data_ptr(), a.data_ptr(), b.data_ptr(), B, Bg, D, 0, 1.0, 1e-12,
              pack.data_ptr(), n2.data_ptr(), pack.data_ptr() + 8 * B, 0, 0, 0, 0.0, 0.0, 0, ws.data_ptr(), nb, 2, 0,
              _lib.stream_ptr())
    _lib.call("cfa_sparc_finalize", pack.data_ptr(), Bg, part.data_ptr(), m.view(torch.uint8).data_ptr(), B, T, 1.0, 1.0,
              out8.data_ptr(), N, _lib.stream_ptr())
    _lib.call("cfa_global_infonce_bwd", a.data_ptr(), b.data_ptr(), a.data_ptr(), b.data_ptr(), B, Bg, D, 0, 1.0, 1e-12,
              pack.data_ptr(), pack.data_ptr(), n2.data_ptr(), coef.data_ptr(), dab.data_ptr(), dab[1].data_ptr(),
              ws.data_ptr(), nb, 2, 0, _lib.stream_ptr())
    torch.cuda.synchronize()


def _gathered_oracle(vs, ls, mask, thr, s, semantics="reference"):
    """fp64 oracle of N ranks' total_loss (gw = lw = 1) and of d total_r / d (v_r, l_r) with all-gather-with-grad
    semantics for the global InfoNCE (gradient of the GLOBAL mean loss w.r.t. the local rows)."""
    N, B, P = len(vs), vs[0].shape[0], vs[0].shape[1]
    fw = [lo.sparc_forward(v.double(), l.double(), mask, thr, 0.0, 1.0, s, mask_semantics=semantics) for v, l in zip(vs, ls)]
    grads = [lo.sparc_backward(f) for f in fw]                                   # fine-grained part only (gw = 0)
    a = torch.cat([f["_cache"]["vbar_raw"] for f in fw])                         # pooled image rows of all ranks
    b = torch.cat([f["_cache"]["lbar_raw"] for f in fw])
    f1 = lo.infonce_forward(a, b, s)
    f2 = lo.infonce_forward(b, a, s)
    Bg = N * B
    glob = 0.5 * (f1["loss_sum"] + f2["loss_sum"]) / Bg
    da, db = lo.symmetric_infonce_backward(f1["ah"], f1["an"], f1["bh"], f1["bn"], f1["lse"], f2["lse"], s, 0.5, 0.5, float(Bg))
    out = []
    for r in range(N):
        sl = slice(r * B, (r + 1) * B)
        cnt = fw[r]["_cache"]["cnt"]
        dv = grads[r][0] + (da[sl] / P)[:, None, :]
        dl = grads[r][1] + (db[sl] / cnt)[:, None, :] * mask[..., None].double()
        out.append(dict(total=float(glob + fw[r]["local_loss"]), glob=float(glob), local=float(fw[r]["local_loss"]),
                        vl=float(f1["loss_sum"] / Bg), lv=float(f2["loss_sum"] / Bg), dv=dv, dl=dl))
    return out


# last two cases: padded masks ("truncate" semantics) and fp16 embeddings (range-scaled fp16 operands on the
# tensor cores)
@pytest.mark.parametrize("N,B,P,T,D,s,dtype,padded", [(2, 64, 50, 20, 256, 1.0, torch.bfloat16, False),
                                                      (4, 24, 196, 77, 512, 2.0, torch.bfloat16, False),
                                                      (3, 130, 33, 20, 256, 1.0, torch.bfloat16, False),
                                                      (2, 40, 50, 20, 256, 1.0, torch.bfloat16, True),
                                                      (2, 24, 50, 20, 128, 1.0, torch.float16, False)])
def test_gathered_sparc_loss_peer_memory_emulated_ranks(N, B, P, T, D, s, dtype, padded):
    from clip_finegrained_alignment_b200 import _lib
    L = _lib.lib
    g = torch.Generator().manual_seed(1000 * N + B)
    vs = [torch.randn(B, P, D, generator=g).to(dtype) for _ in range(N)]
    ls = [torch.randn(B, T, D, generator=g).to(dtype) for _ in range(N)]
    mask = torch.ones(B, T, dtype=torch.bool)
    if padded:
        for b in range(0, B, 3):
            mask[b, T - 1 - (b % 5):] = False
    thr = float(torch.tensor(1.0 / P, dtype=torch.float32))
    ref = _gathered_oracle([v.float() for v in vs], [l.float() for l in ls], mask, thr, s,
                           "truncate" if padded else "reference")

    _load_kernels(B, P, T, D, N, dtype)
    ex = _LocalExchange(N, B, D)
    try:
        code = _lib.DTYPE_CODE[dtype]
        nbytes = L.cfa_sparc_loss_gathered_workspace_bytes(B, P, T, D, code, 0, N)
        assert nbytes > L.cfa_sparc_loss_workspace_bytes(B, P, T, D, code, 0)
        streams = [torch.cuda.Stream() for _ in range(N)]
        vc = [v.cuda() for v in vs]
        lc = [l.cuda() for l in ls]
        mu8 = mask.cuda().view(torch.uint8)
        one = torch.ones((), device="cuda")
        for step in range(3):                                   # both slots and repeated epochs
            ws = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(N)]
            dv = [torch.empty_like(x) for x in vc]
            dl = [torch.empty_like(x) for x in lc]
            torch.cuda.synchronize()
            for r in range(N):
                with torch.cuda.stream(streams[r]):
                    _lib.call("cfa_sparc_loss_gathered_fwd", vc[r].data_ptr(), lc[r].data_ptr(), mu8.data_ptr(), B, P, T, D,
                              code, thr, s, 1.0, 1.0, ws[r].data_ptr(), nbytes, 0, N, r, ex.blocks, step, _lib.stream_ptr())
            for r in range(N):                                  # the backward needs no exchange: any order, any time
                with torch.cuda.stream(streams[r]):
                    _lib.call("cfa_sparc_loss_gathered_bwd", vc[r].data_ptr(), lc[r].data_ptr(), mu8.data_ptr(), B, P, T, D,
                              code, thr, s, 1.0, 1.0, ws[r].data_ptr(), nbytes, 0, 0, one.data_ptr(), 0, 0, 0, 0,
                              dv[r].data_ptr(), dl[r].data_ptr(), 0, N, r, _lib.stream_ptr())
            torch.cuda.synchronize()
            for r in range(N):
                out = ws[r][:32].view(torch.float32).cpu()
                o = ref[r]
                for got, want, name in ((out[0], o["glob"], "global"), (out[1], o["local"], "local"), (out[2], o["total"], "total"),
                                        (out[3], o["vl"], "vl"), (out[4], o["lv"], "lv")):
                    assert abs(float(got) - want) <= 1e-4 * abs(want), (step, r, name, float(got), want)
                # gradients: 1e-3 + one bf16 output rounding (2^-8), as for the rank-local tensor-core path
                assert rel_err(dv[r].float(), o["dv"]) <= 1e-3 + 2 ** -8, (step, r, rel_err(dv[r].float(), o["dv"]))
                assert rel_err(dl[r].float(), o["dl"]) <= 1e-3 + 2 ** -8, (step, r, rel_err(dl[r].float(), o["dl"]))
    finally:
        ex.close()


@pytest.mark.parametrize("N,B,D,temp", [(3, 70, 256, 0.07), (2, 256, 512, 1.0)])
def test_gathered_clip_loss_peer_memory_emulated_ranks(N, B, D, temp):
    """cfa_global_infonce_gathered_fwd/_bwd (CustomCLIPLoss with gather=True): N ranks in one process, oracle =
    CustomCLIPLoss restated on the concatenated batch (losses.py:14-36)."""
    from clip_finegrained_alignment_b200 import _lib
    L = _lib.lib
    g = torch.Generator().manual_seed(17 * N + B)
    img = torch.randn(N * B, D, generator=g); txt = torch.randn(N * B, D, generator=g)
    fw = lo.clip_loss_forward(img.double(), txt.double(), temp)
    da_ref, db_ref = lo.clip_loss_backward(fw, temp)
    Bg = N * B
    # load every kernel of the path first (see _load_kernels): tensor-core global forward + backward, rank-locally
    a = img.cuda(); b = txt.cuda()
    nb = L.cfa_global_infonce_workspace_bytes(B, Bg, D)
    w0 = torch.empty(nb, dtype=torch.uint8, device="cuda")
    pk = torch.zeros(N, 2 * B + 2, device="cuda"); n2 = torch.empty(2, B, device="cuda"); cf = torch.zeros(2, device="cuda")
    d0 = torch.empty(2, B, D, device="cuda")
    _lib.call("cfa_global_infonce_fwd", a.data_ptr(), b.data_ptr(), a.data_ptr(), b.data_ptr(), B, Bg, D, 0, 1.0, 0.0,
              pk.data_ptr(), n2.data_ptr(), pk.data_ptr() + 8 * B, 0, 0, 0, 0.0, 0.0, 0, w0.data_ptr(), nb, 2, 0, _lib.stream_ptr())
    _lib.call("cfa_global_infonce_bwd", a.data_ptr(), b.data_ptr(), a.data_ptr(), b.data_ptr(), B, Bg, D, 0, 1.0, 0.0,
              pk.data_ptr(), pk.data_ptr(), n2.data_ptr(), cf.data_ptr(), d0.data_ptr(), d0[1].data_ptr(), w0.data_ptr(), nb, 2, 0,
              _lib.stream_ptr())
    blocks1 = _LocalExchange(1, B, D)
    _lib.call("cfa_peer_sync", blocks1.blocks, 1, 0, 1, n2.data_ptr(), 64, 8, 64, 8, d0.data_ptr(), _lib.stream_ptr())
    torch.cuda.synchronize()
    blocks1.close()

    ex = _LocalExchange(N, B, D)
    try:
        nbytes = L.cfa_global_infonce_gathered_workspace_bytes(B, D, N)
        streams = [torch.cuda.Stream() for _ in range(N)]
        ab = [torch.stack([img[r * B:(r + 1) * B], txt[r * B:(r + 1) * B]]).cuda().contiguous() for r in range(N)]
        coef = torch.full((2,), 0.5 / Bg, device="cuda")
        for step in range(2):
            ws = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(N)]
            sums = [torch.zeros(2, device="cuda") for _ in range(N)]
            dab = [torch.empty_like(x) for x in ab]
            torch.cuda.synchronize()
            for r in range(N):
                with torch.cuda.stream(streams[r]):
                    _lib.call("cfa_global_infonce_gathered_fwd", ab[r].data_ptr(), B, D, 1.0 / temp, 0.0, ws[r].data_ptr(), nbytes,
                              N, r, ex.blocks, step, sums[r].data_ptr(), _lib.stream_ptr())
            for r in range(N):
                with torch.cuda.stream(streams[r]):
                    _lib.call("cfa_global_infonce_gathered_bwd", ab[r].data_ptr(), B, D, 1.0 / temp, 0.0, ws[r].data_ptr(), nbytes,
                              coef.data_ptr(), dab[r].data_ptr(), N, r, _lib.stream_ptr())
            torch.cuda.synchronize()
            want = float(fw["clip_loss"])
            for r in range(N):
                got = float(sums[r].sum()) * 0.5 / Bg
                assert abs(got - want) <= 2e-5 * max(1.0, abs(want)), (step, r, got, want)
                sl = slice(r * B, (r + 1) * B)
                assert rel_err(dab[r][0], da_ref[sl]) <= 2e-4, (step, r, rel_err(dab[r][0], da_ref[sl]))
                assert rel_err(dab[r][1], db_ref[sl]) <= 2e-4, (step, r, rel_err(dab[r][1], db_ref[sl]))
    finally:
        ex.close()


def test_gathered_unsupported_shapes_are_refused():
    """fp32 inputs / D the tensor-core global kernels do not take: CFA_ERR_UNSUPPORTED (-2), caller keeps NCCL."""
    from clip_finegrained_alignment_b200 import _lib
    L = _lib.lib
    blocks = (C.c_void_p * 2)(1, 1)
    ws = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
    x = torch.zeros(8, device="cuda")
    rc = L.cfa_sparc_loss_gathered_fwd(x.data_ptr(), x.data_ptr(), x.data_ptr(), 2, 5, 3, 96, 0, 0.2, 1.0, 1.0, 1.0,
                                       ws.data_ptr(), ws.numel(), 0, 2, 0, blocks, 0, 0)
    assert rc == -2
    rc = L.cfa_sparc_loss_gathered_fwd(x.data_ptr(), x.data_ptr(), x.data_ptr(), 2, 5, 3, 96, 0, 0.2, 1.0, 1.0, 1.0,
                                       ws.data_ptr(), ws.numel(), 0, 1, 0, blocks, 0, 0)
    assert rc == -1                                            # world < 2


_WORKER = r'''
import os, sys, json, types
sys.path.insert(0, os.environ["CFA_ROOT"])
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
from clip_finegrained_alignment_b200 import SPARCLoss
B, P, T, D = 96, 196, 77, 512
cfg = types.SimpleNamespace(similarity_threshold=1.0 / P, global_loss_weight=1.0, local_loss_weight=1.0, inverse_temperature=1.0)
torch.manual_seed(42 + rank)
v0 = torch.randn(B, P, D, device="cuda").to(torch.bfloat16)
l0 = torch.randn(B, T, D, device="cuda").to(torch.bfloat16)
m = torch.ones(B, T, dtype=torch.bool, device="cuda")
res = {}
for mode in (True, "nccl"):
    crit = SPARCLoss(cfg, gather=mode)
    for it in range(3):
        v = v0.clone().requires_grad_(True); l = l0.clone().requires_grad_(True)
        out = crit(v, l, m)
        out["total_loss"].backward()
    torch.cuda.synchronize()
    res[str(mode)] = (out["total_loss"].item(), out["global_loss"].item(), v.grad.float().cpu(), l.grad.float().cpu())
from clip_finegrained_alignment_b200 import CustomCLIPLoss, peer
used_peer = any(e.ok for e in peer._EXCHANGES.values())
clip = {}
i0 = torch.randn(B, D, device="cuda").to(torch.bfloat16); t0 = torch.randn(B, D, device="cuda").to(torch.bfloat16)
for mode in (True, "nccl"):
    crit = CustomCLIPLoss(0.07, gather=mode)
    for it in range(2):
        i1 = i0.clone().requires_grad_(True); t1 = t0.clone().requires_grad_(True)
        o = crit(i1, t1)
        o["total_loss"].backward()
    torch.cuda.synchronize()
    clip[str(mode)] = (o["clip_loss"].item(), i1.grad.float().cpu(), t1.grad.float().cpu())
ca, cb = clip["True"], clip["nccl"]
clip_ok = abs(ca[0] - cb[0]) <= 1e-6 * abs(cb[0]) and float((ca[1] - cb[1]).norm() / cb[1].norm()) <= 1e-4 \
    and float((ca[2] - cb[2]).norm() / cb[2].norm()) <= 1e-4
a, b = res["True"], res["nccl"]
ok = used_peer and clip_ok and abs(a[0] - b[0]) <= 1e-6 * abs(b[0]) and abs(a[1] - b[1]) <= 1e-6 * abs(b[1]) \
    and float((a[2] - b[2]).norm() / b[2].norm()) <= 1e-4 and float((a[3] - b[3]).norm() / b[3].norm()) <= 1e-4
# ---- against the ORACLE on the concatenated batch (not only against the NCCL path of the same kernels)
from oracle import losses_oracle as lo
vs = [torch.empty_like(v0) for _ in range(world)]; ls = [torch.empty_like(l0) for _ in range(world)]
dist.all_gather(vs, v0); dist.all_gather(ls, l0)
thr = float(torch.tensor(1.0 / P, dtype=torch.float32))
mc = m.cpu()
fw = [lo.sparc_forward(x.cpu().double(), y.cpu().double(), mc, thr, 0.0, 1.0, 1.0) for x, y in zip(vs, ls)]
ga, gb = torch.cat([f["_cache"]["vbar_raw"] for f in fw]), torch.cat([f["_cache"]["lbar_raw"] for f in fw])
f1, f2 = lo.infonce_forward(ga, gb, 1.0), lo.infonce_forward(gb, ga, 1.0)
Bg = world * B
glob = float(0.5 * (f1["loss_sum"] + f2["loss_sum"]) / Bg)
da, db = lo.symmetric_infonce_backward(f1["ah"], f1["an"], f1["bh"], f1["bn"], f1["lse"], f2["lse"], 1.0, 0.5, 0.5, float(Bg))
gv, gl = lo.sparc_backward(fw[rank])
sl = slice(rank * B, (rank + 1) * B)
# default gather_grad_reduce="mean": the global term carries the factor `world` (DDP averages the ranks afterwards)
dv_ref = gv + world * (da[sl] / P)[:, None, :]
dl_ref = gl + world * (db[sl] / fw[rank]["_cache"]["cnt"])[:, None, :] * mc[..., None].double()
rel = lambda x, r: float((x.double() - r).norm() / r.norm())
oracle_ok = abs(a[1] - glob) <= 1e-4 * abs(glob) and abs(a[0] - (glob + float(fw[rank]["local_loss"]))) <= 1e-4 * abs(a[0]) \
    and rel(a[2], dv_ref) <= 1e-3 + 2 ** -8 and rel(a[3], dl_ref) <= 1e-3 + 2 ** -8
ok = ok and oracle_ok
print(json.dumps({"rank": rank, "ok": bool(ok), "oracle_ok": bool(oracle_ok), "used_peer": bool(used_peer), "total": a[0],
                  "total_nccl": b[0], "dv_rel": rel(a[2], dv_ref), "dl_rel": rel(a[3], dl_ref)}), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


@pytest.mark.timeout(600)
def test_gathered_sparc_loss_two_processes_cuda_ipc(tmp_path):
    """Two processes, two GPUs: SPARCLoss(gather=True) over CUDA-IPC peer memory == SPARCLoss(gather='nccl')
    (same kernels behind both; the exchange is the only difference)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "peer_worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, CFA_ROOT=ROOT, NCCL_DEBUG="WARN")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29731", str(script)],
                       env=env, capture_output=True, text=True, timeout=540)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count('"ok": true') == 2, r.stdout[-3000:]
