import sys, ctypes, torch
sys.path.insert(0, '.')
from clip_finegrained_alignment_b200 import _lib
L = _lib.lib
N, B, D, s = (int(sys.argv[1]) if len(sys.argv) > 1 else 2), (int(sys.argv[2]) if len(sys.argv) > 2 else 256), 512, 1.0
Bg = N * B
mark = torch.zeros(2048 * 16 * 32, dtype=torch.int32).pin_memory()
a = torch.randn(Bg, D).cuda(); b = torch.randn(Bg, D).cuda()
ws_bytes = L.cfa_global_infonce_workspace_bytes(B, Bg, D)
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
l2 = torch.empty(2, B, device="cuda"); n2 = torch.empty(2, B, device="cuda"); s2 = torch.empty(2, device="cuda")
lall = torch.zeros(2, Bg, device="cuda")
coef = torch.full((2,), 0.5 / Bg, device="cuda")
da = torch.empty(B, D, device="cuda"); db = torch.empty(B, D, device="cuda")
for it in range(3):
    if it == 2: L.cfa_debug_set_marker_buffer(mark.data_ptr())
    _lib.call("cfa_global_infonce_fwd", a.data_ptr(), b.data_ptr(), a.data_ptr(), b.data_ptr(), B, Bg, D, 0, s, 1e-12,
              l2.data_ptr(), n2.data_ptr(), s2.data_ptr(), 0, 0, 0, 0.0, 0.0, 0, ws.data_ptr(), ws_bytes, 2, 0, _lib.stream_ptr())
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call("cfa_global_infonce_bwd", a.data_ptr(), b.data_ptr(), a.data_ptr(), b.data_ptr(), B, Bg, D, 0, s, 1e-12,
              l2.data_ptr(), lall.data_ptr(), n2.data_ptr(), coef.data_ptr(), da.data_ptr(), db.data_ptr(), ws.data_ptr(), ws_bytes, 2, 0, _lib.stream_ptr())
    e1.record(); torch.cuda.synchronize()
    print('bwd call ms', e0.elapsed_time(e1))
m = mark.view(-1, 16, 32)
names = {1:'start',2:'alloc done',3:'sync1',4:'phaseA issued',5:'ds_ready seen',6:'phaseC issued',10:'s_full seen',11:'ld done',12:'bar',13:'dS written',14:'o_done seen',15:'stores done',20:'end'}
for cta in (0, 5, 700):
    print('CTA', cta)
    for w in range(10):
        print('  warp', w, {names[k]: int(m[cta, w, k]) for k in names if int(m[cta, w, k]) != 0})
