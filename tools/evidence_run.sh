set -x
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest12.log 2>&1; tail -n 2 gpurun_out/pytest12.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke3.log 2>&1; tail -n 1 gpurun_out/smoke3.log
timeout 300 python bench.py > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err; echo bench rc=$?
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1i_ref.json 2> gpurun_out/bench_r1i_ref.err; echo ref rc=$?
timeout 300 python bench.py --batch 1024 --no-adamspd --no-cpu-baseline > gpurun_out/bench_r1i_b1024.json 2> gpurun_out/bench_r1i_b1024.err; echo b1024 rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1i.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-adamspd > gpurun_out/ncu9.log 2>&1; echo ncu1 rc=$?
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'sparc_bwd2_kernel|sparc_fwd2_kernel' -s 6 -c 2 -o gpurun_out/prof_r1i -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-adamspd > gpurun_out/ncu10.log 2>&1; echo ncu2 rc=$?
timeout 300 ncu --set full --clock-control none -k regex:'adamspd_gradnorm|adamspd_pass1' -c 2 -o gpurun_out/prof_r1i_amp -f python -c "
import torch, sys
sys.path.insert(0, '.')
from clip_finegrained_alignment_b200 import AdamSPD
ps = [torch.nn.Parameter(torch.randn(4096, 4096, device='cuda') * 0.02) for _ in range(8)]
pre = [p.detach() + 1e-3 for p in ps]
opt = AdamSPD([{'params': ps, 'pre': pre}], lr=2e-5, weight_decay=0.1)
for p in ps: p.grad = torch.randn_like(p) * 1e-3
opt.amp_step(None, 1.0); torch.cuda.synchronize()
" > gpurun_out/ncu11.log 2>&1; echo ncu3 rc=$?
