# tools/run_r2b.sh [full]: GPU tests + short benches of every config after a kernel change (one gpurun call)
if [ "$1" = full ]; then python -m pytest tests -x -q -m gpu > gpurun_out/t_r2b.log 2>&1; else
python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py -x -q -m gpu > gpurun_out/t_r2b.log 2>&1; fi; tail -4 gpurun_out/t_r2b.log
for c in cfg2 cfg4 cfg3 cfg1 cfg5; do
  python bench.py --config $c --steps 3 --warmup 3 --no-cpu > gpurun_out/b_r2b_$c.log 2>&1
  tail -1 gpurun_out/b_r2b_$c.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$c', round(d['value']/1e6,2), round(d['e2e']['value']/1e6,2), round(d['roofline']['frac'],3), d['e2e']['records_identical_to_resident_run'], round(d['roofline']['candidates_per_unit'],1))" || tail -5 gpurun_out/b_r2b_$c.log
done
