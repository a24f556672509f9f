# tools/final_measure.sh: everything the round's numbers come from, on one GPU box (writes under gpurun_out/)
set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-200
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-200
python bench_configs.py > gpurun_out/bench_configs.log 2>&1; grep -c "^{" gpurun_out/bench_configs.log
