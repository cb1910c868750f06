"""2-rank NCCL check: SPARCLoss(gather=True) / CustomCLIPLoss(gather=True) vs the oracle on the concatenated batch."""
import os, sys, types
import torch, torch.distributed as dist
sys.path.insert(0, '.')
from clip_finegrained_alignment_b200 import SPARCLoss, CustomCLIPLoss
from oracle import losses_oracle as lo
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr_ = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
B, P, T, D, s = 6, 50, 77, 256, 2.0
g = torch.Generator().manual_seed(5)
V = torch.randn(world * B, P, D, generator=g).to(torch.bfloat16)
L = torch.randn(world * B, T, D, generator=g).to(torch.bfloat16)
M = torch.ones(world * B, T, dtype=torch.bool)
cfg = types.SimpleNamespace(similarity_threshold=1.0 / P, global_loss_weight=1.0, local_loss_weight=1.0, inverse_temperature=s)
sl = slice(rank * B, (rank + 1) * B)
v = V[sl].cuda().requires_grad_(True); l = L[sl].cuda().requires_grad_(True)
out = SPARCLoss(cfg, gather=True)(v, l, M[sl].cuda())
out["global_loss"].backward()
# oracle: global loss on the concatenated batch (pooled embeddings), gradient w.r.t. this rank's rows
Vd, Ld = V.double().requires_grad_(True), L.double().requires_grad_(True)
vb = torch.nn.functional.normalize(Vd.mean(1), dim=-1); lb = torch.nn.functional.normalize(Ld.mean(1), dim=-1)
f1 = lo.infonce_forward(vb, lb, s); f2 = lo.infonce_forward(lb, vb, s)
ref = 0.5 * (f1["loss_sum"] + f2["loss_sum"]) / (world * B)
ref.backward()
e_loss = abs(float(out["global_loss"]) - float(ref))
e_dv = float((v.grad.double().cpu() - Vd.grad[sl]).norm() / Vd.grad[sl].norm())
e_dl = float((l.grad.double().cpu() - Ld.grad[sl]).norm() / Ld.grad[sl].norm())
# CLIP loss gathered
a = torch.randn(world * 8, 64, generator=g); b = torch.randn(world * 8, 64, generator=g)
aa = a[rank * 8:(rank + 1) * 8].cuda().requires_grad_(True); bb = b[rank * 8:(rank + 1) * 8].cuda().requires_grad_(True)
o2 = CustomCLIPLoss(0.07, gather=True)(aa, bb); o2["total_loss"].backward()
oc = lo.clip_loss_forward(a.double(), b.double(), 0.07); da, db = lo.clip_loss_backward(oc, 0.07)
e_c = abs(float(o2["clip_loss"]) - float(oc["clip_loss"]))
e_da = float((aa.grad.double().cpu() - da[rank * 8:(rank + 1) * 8]).norm() / da[rank * 8:(rank + 1) * 8].norm())
print(f"rank {rank}: sparc global loss err {e_loss:.2e} dv {e_dv:.2e} dl {e_dl:.2e} | clip loss err {e_c:.2e} da {e_da:.2e}", flush=True)
assert e_loss < 1e-4 and e_dv < 1e-2 and e_dl < 1e-2 and e_c < 1e-4 and e_da < 1e-4
dist.destroy_process_group()
