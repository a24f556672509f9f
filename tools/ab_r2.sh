python -m pytest tests -x -q -m gpu > gpurun_out/t_r2b.log 2>&1; tail -3 gpurun_out/t_r2b.log
bash tools/ncu_kernel.sh r2g_wide cfg5 bsx_map_se_wide 200000 3
