import os, sys, time, torch
sys.path.insert(0, '.')
import bench
dev = torch.device('cuda:0')
print('affinity', len(os.sched_getaffinity(0)), 'local cpus', bench.gpu_local_cpus(dev))
try:
    print(open('/sys/devices/system/node/online').read().strip(), [open(f'/sys/devices/system/node/node{n}/cpulist').read().strip() for n in range(8) if os.path.exists(f'/sys/devices/system/node/node{n}')])
except Exception as e: print(e)
def bw(cpus):
    prev = os.sched_getaffinity(0)
    if cpus: os.sched_setaffinity(0, cpus)
    h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); h.fill_(1)
    os.sched_setaffinity(0, prev)
    d = torch.empty_like(h, device=dev)
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return 5 * h.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9
allc = sorted(os.sched_getaffinity(0))
loc = bench.gpu_local_cpus(dev)
print('H2D GB/s default', bw(None))
if loc:
    print('H2D GB/s local ', bw(loc))
    rem = [c for c in allc if c not in loc]
    if rem: print('H2D GB/s remote', bw(rem))
