# full GPU suite, then ncu captures of the five hot kernels (summaries go to profiles/ from gpurun_out/)
python -m pytest tests -x -q -m gpu > gpurun_out/t_r2c.log 2>&1; tail -4 gpurun_out/t_r2c.log
bash tools/ncu_kernel.sh r2f_se cfg2 bsx_map_se_wgbs 4000000 3
bash tools/ncu_kernel.sh r2f_pe cfg3 bsx_map_pe 2000000 3
bash tools/ncu_kernel.sh r2f_rrbs cfg4 bsx_map_se_rrbs 2000000 3
bash tools/ncu_kernel.sh r2f_wide cfg5 bsx_map_se_wide 200000 3
