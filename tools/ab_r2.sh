python -m pytest tests -x -q -m gpu > gpurun_out/t_r2b.log 2>&1; tail -2 gpurun_out/t_r2b.log
for c in cfg5 cfg4; do
  python bench.py --config $c --steps 3 --warmup 3 --no-cpu > gpurun_out/b_r2b_$c.log 2>&1
  tail -1 gpurun_out/b_r2b_$c.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$c', round(d['value']/1e6,2), round(d['e2e']['value']/1e6,2), round(d['roofline']['frac'],3), d['e2e']['records_identical_to_resident_run'], round(d['roofline']['candidates_per_unit'],1))" || tail -5 gpurun_out/b_r2b_$c.log
done
