# End-of-round single-GPU evidence run (under gpurun): GPU tests, smoke, bench (+ fp16, reference arm), ncu launch list of
# the bench command, one --set full capture of the two SPARC kernels at the bench batch, phase stamps.  TAG = file prefix.
TAG=${1:-r2n}
set -x
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest.log 2>&1; tail -n 2 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -n 1 gpurun_out/${TAG}_smoke.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo bench rc=$?
timeout 300 python bench.py --dtype f16 --no-adamspd --no-cpu > gpurun_out/${TAG}_bench_f16.json 2>/dev/null; echo f16 rc=$?
timeout 300 python bench.py --no-graph --no-adamspd --no-cpu > gpurun_out/${TAG}_bench_eager.json 2>/dev/null; echo eager rc=$?
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2>/dev/null; echo ref rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-adamspd --no-cpu > gpurun_out/${TAG}_ncu1.log 2>&1; echo ncu1 rc=$?
python tools/run_once.py 256 && timeout 400 ncu --set full --import-source on --clock-control none -k regex:'sparc_(fwd3|bwd3)' --launch-skip 4 -c 2 -o gpurun_out/${TAG}_gen3_B256 -f python tools/run_once.py 256 > gpurun_out/${TAG}_ncu2.log 2>&1; echo ncu2 rc=$?
python tools/prof_gen3.py 148 2>&1 | grep -v Warn > gpurun_out/${TAG}_phase_stamps_B148_warm.txt
python tools/prof_gen3.py 256 cold 2>&1 | grep -v Warn > gpurun_out/${TAG}_phase_stamps_B256_L2cold.txt
python tools/prof_global8.py 2>&1 | grep -v Warn > gpurun_out/${TAG}_global_config3_kernels.txt
