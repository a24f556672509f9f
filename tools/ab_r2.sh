python -m pytest tests -x -q -m gpu > gpurun_out/t_final.log 2>&1; tail -2 gpurun_out/t_final.log
bash tools/sanitize_r2.sh
