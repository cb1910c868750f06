"""Per-kernel device times of the gathered global InfoNCE at BASELINE config 3 scale on ONE GPU
(B = 1024 local rows x Bg = 8192 global columns, D = 512: what each of 8 ranks computes)."""
import sys, torch
sys.path.insert(0, '.')
from clip_finegrained_alignment_b200 import _lib
from torch.profiler import profile, ProfilerActivity
L = _lib.lib
B, Bg, D, s = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 8192, 512, 1.0
a = torch.randn(Bg, D).cuda(); b = torch.randn(Bg, D).cuda()
ws_bytes = L.cfa_global_infonce_workspace_bytes(B, Bg, D)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
l2 = torch.empty(2, B, device="cuda"); n2 = torch.empty(2, B, device="cuda"); s2 = torch.empty(2, device="cuda")
lall = torch.zeros(2, Bg, device="cuda")
coef = torch.full((2,), 0.5 / Bg, device="cuda")
da = torch.empty(B, D, device="cuda"); db = torch.empty(B, D, device="cuda")
def run():
    _lib.call("cfa_global_infonce_fwd", a.data_ptr(), b.data_ptr(), a.data_ptr(), b.data_ptr(), B, Bg, D, 0, s, 1e-12,
              l2.data_ptr(), n2.data_ptr(), s2.data_ptr(), 0, 0, 0, 0.0, 0.0, 0, ws.data_ptr(), ws_bytes, 2, 0, _lib.stream_ptr())
    _lib.call("cfa_global_infonce_bwd", a.data_ptr(), b.data_ptr(), a.data_ptr(), b.data_ptr(), B, Bg, D, 0, s, 1e-12,
              l2.data_ptr(), lall.data_ptr(), n2.data_ptr(), coef.data_ptr(), da.data_ptr(), db.data_ptr(), ws.data_ptr(), ws_bytes, 2, 0, _lib.stream_ptr())
for _ in range(3): run()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(10): run()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / max(1, e.count), e.count) for e in prof.key_averages() if e.device_time_total > 0]
print(f"B={B} Bg={Bg} D={D}")
for k, t, c in sorted(rows, key=lambda r: -r[1] * r[2])[:12]:
    print(f'{k[:90]:90s} {t:9.1f} us x {c}')
