python -m pytest tests/test_gpu_parity.py tests/test_scale_gpu.py -x -q -m gpu > gpurun_out/t_r2b.log 2>&1; tail -3 gpurun_out/t_r2b.log
python tests/cli_bench.py --reads 20000000 --len 100 --genome-mb 3100 --opts "-s 16 -v 5 -I 4 -S 7" --skip-ref --repeat 2 > gpurun_out/cli_r2_cfg2.json 2> gpurun_out/cli_r2_cfg2.err; tail -c 900 gpurun_out/cli_r2_cfg2.json
