import os, sys, types, time, torch, torch.distributed as dist
sys.path.insert(0, '.')
from clip_finegrained_alignment_b200 import SPARCLoss, _lib
rank = int(os.environ["RANK"]); lr_ = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))
def cfg(thr, s=1.0): return types.SimpleNamespace(similarity_threshold=thr, global_loss_weight=1.0, local_loss_weight=1.0, inverse_temperature=s)
B,P,T,D = 256,196,77,512
torch.manual_seed(rank)
v = torch.randn(B,P,D,device='cuda').to(torch.bfloat16).requires_grad_(True)
l = torch.randn(B,T,D,device='cuda').to(torch.bfloat16).requires_grad_(True)
m = torch.ones(B,T,dtype=torch.bool,device='cuda')
crit = SPARCLoss(cfg(1.0/P), gather=True)
def step():
    v.grad=None; l.grad=None
    crit(v,l,m)['total_loss'].backward()
for _ in range(5): step()
torch.cuda.synchronize(); dist.barrier()
N=50
t0=time.perf_counter()
for _ in range(N): step()
t1=time.perf_counter()
torch.cuda.synchronize()
t2=time.perf_counter()
if rank==0: print(f'host issue time/step {1e6*(t1-t0)/N:.0f} us ; wall/step incl. drain {1e6*(t2-t0)/N:.0f} us')
_lib.kernel_events = {k: [] for k in _lib.LAUNCHES}
for _ in range(10): step()
torch.cuda.synchronize()
if rank==0:
    for k,ev in _lib.kernel_events.items():
        if ev: print(f'  {k:28s} {1e3*sum(a.elapsed_time(b) for a,b in ev)/10:8.1f} us/step')
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
if rank==0: print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=50))
dist.destroy_process_group()
