import sys, torch
sys.path.insert(0,'.')
import bench
pk = bench.peaks()
out = bench.bench_adamspd(torch.device('cuda',0), 4, 2, pk)
print(out)
