BSMAP_B200_LIB=variants/v3.so python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py -x -q -m gpu -k "not cli" > gpurun_out/t_v3.log 2>&1; tail -2 gpurun_out/t_v3.log
bash tools/ab_bench.sh v3
CFG=cfg4 bash tools/ab_bench.sh v3
