BSMAP_B200_LIB=variants/pipe4.so python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py -x -q -m gpu -k "not cli" > gpurun_out/t_pipe.log 2>&1; tail -2 gpurun_out/t_pipe.log
bash tools/ab_bench.sh se_pfc pipe4
CFG=cfg3 bash tools/ab_bench.sh pipe4
