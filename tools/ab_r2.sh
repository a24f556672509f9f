python -m pytest tests/test_gpu_parity.py tests/test_methratio_gpu.py tests/test_glue_gpu.py -x -q -m gpu -k "cli or glue" > gpurun_out/t_cli.log 2>&1; tail -2 gpurun_out/t_cli.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launch_r2_final.log 2>&1; tail -1 gpurun_out/launch_r2_final.log | cut -c1-120
bash tools/cli_r2.sh
