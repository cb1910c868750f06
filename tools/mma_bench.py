import sys, torch
sys.path.insert(0,'.')
from clip_finegrained_alignment_b200 import _lib
names_a={0:'A sw128 K',1:'A il K',2:'A il MN',3:'A sw128 MN',4:'A sw64 K'}
names_b={0:'B sw128 K',1:'B sw128 MN',2:'B il K',3:'B il MN',4:'B sw64 K',5:'B sw64 MN'}
cases=[(0,0,208,128),(0,0,64,128),(0,0,32,128),(4,4,208,64),(4,4,32,64),(1,0,208,128),(1,0,64,128),(1,1,64,208),(1,5,32,208),(1,5,32,80),(2,5,32,80),(2,3,32,80),(1,3,64,208),(1,4,208,32),(1,2,96,64),(3,3,160,208),(3,3,80,208),(3,2,208,80),(2,3,160,208),(2,3,80,208),(0,0,240,128),(0,0,80,128),(1,2,80,80),(1,3,80,80)]
cyc=torch.zeros(1,dtype=torch.int64,device='cuda')
for a,b,N,K in cases:
    A=torch.randn(128 if a not in (2,3) else K, K if a not in (2,3) else 128).to(torch.bfloat16).cuda()
    Bm=torch.randn(N,K).to(torch.bfloat16)
    Bm=(Bm.t().contiguous() if b in (1,3,5) else Bm).cuda()
    D=torch.empty(128,N,device='cuda')
    res=[]
    for rep in (1,64):
        _lib.call("cfa_tc_selftest_timed",a,b,N,K,A.data_ptr(),Bm.data_ptr(),D.data_ptr(),rep,cyc.data_ptr(),_lib.stream_ptr())
        torch.cuda.synchronize(); res.append(int(cyc.item()))
    n1=K//16; n64=64*n1
    per=(res[1]-res[0])/(n64-n1)
    print(f'{names_a[a]:10s} x {names_b[b]:11s} N={N:3d} K={K:3d}: {per:7.1f} cycles/MMA  (ideal {128*N/256:.0f})')
