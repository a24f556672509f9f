nvidia-smi -L | wc -l
N=${NG:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_r2_${N}gpu.log 2>&1; tail -1 gpurun_out/bench_r2_${N}gpu.log | cut -c1-300
