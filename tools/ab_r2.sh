for b in 524288 1048576 2097152 4194304; do
python bench.py --steps 3 --warmup 3 --no-cpu --batch $b > gpurun_out/b_batch_$b.log 2>&1
tail -1 gpurun_out/b_batch_$b.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('batch $b', round(d['value']/1e6,2), round(d['e2e']['value']/1e6,2), round(d['e2e']['ascii']['value']/1e6,2))"
done
