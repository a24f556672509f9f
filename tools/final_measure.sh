set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/b_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bsx_map_se -c 1 -s 1 -o gpurun_out/prof_se_final -f python bench.py --steps 1 --warmup 1 --no-cpu --reads 4000000 > gpurun_out/ncu_full.log 2>&1
python bench_configs.py > gpurun_out/bench_configs.log 2>&1; tail -4 gpurun_out/bench_configs.log | cut -c1-250
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
