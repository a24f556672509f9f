python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python -m pytest tests -x -q -m gpu > gpurun_out/t_final.log 2>&1; tail -2 gpurun_out/t_final.log
python bench.py > gpurun_out/bench_r2_default.log 2>&1; tail -1 gpurun_out/bench_r2_default.log | cut -c1-200
for c in cfg3 cfg4; do
  python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/bench_r2_$c.log 2>&1
  tail -1 gpurun_out/bench_r2_$c.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$c', round(d['value']/1e6,2), round(d['e2e']['value']/1e6,2), round(d['roofline']['frac'],3), d['cpu_baseline'] and (round(d['cpu_baseline']['value']), d['cpu_baseline']['records_compared_with_gpu'], d['cpu_baseline']['records_differing_from_gpu']))" || tail -5 gpurun_out/bench_r2_$c.log
done
bash tools/ncu_kernel.sh r2i_se cfg2 bsx_map_se_wgbs 4000000 3
bash tools/ncu_kernel.sh r2i_pe cfg3 bsx_map_pe 2000000 3
