# tools/final_measure.sh: everything the round's single-GPU numbers come from, on one GPU box (writes under gpurun_out/)
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python -m pytest tests -x -q -m gpu > gpurun_out/t_final.log 2>&1; tail -2 gpurun_out/t_final.log
python bench.py > gpurun_out/bench_r2_default.log 2>&1; tail -1 gpurun_out/bench_r2_default.log | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_ref.log 2>&1; tail -1 gpurun_out/bench_r2_ref.log | cut -c1-300
for c in cfg1 cfg3 cfg4 cfg5; do
  python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/bench_r2_$c.log 2>&1
  tail -1 gpurun_out/bench_r2_$c.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$c', round(d['value']/1e6,2), round(d['e2e']['value']/1e6,2), round(d['roofline']['frac'],3), d['cpu_baseline'] and (round(d['cpu_baseline']['value']), d['cpu_baseline']['records_compared_with_gpu'], d['cpu_baseline']['records_differing_from_gpu']))" || tail -5 gpurun_out/bench_r2_$c.log
done
# launch list of the default bench command (cold-cache, serialised: the kernel's share of the step, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launch_r2_final.log 2>&1; tail -1 gpurun_out/launch_r2_final.log | cut -c1-120
bash tools/ncu_kernel.sh r2h_se cfg2 bsx_map_se_wgbs 4000000 3
bash tools/ncu_kernel.sh r2h_pe cfg3 bsx_map_pe 2000000 3
bash tools/ncu_kernel.sh r2h_rrbs cfg4 bsx_map_se_rrbs 2000000 3
bash tools/ncu_kernel.sh r2h_wide cfg5 bsx_map_se_wide 200000 3
bash tools/cli_r2.sh
