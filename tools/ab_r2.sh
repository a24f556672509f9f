python -m pytest tests/test_methratio_gpu.py -x -q -m gpu -k "cli" > gpurun_out/t_meth.log 2>&1; tail -15 gpurun_out/t_meth.log
