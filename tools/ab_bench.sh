# tools/ab_bench.sh [variant ...]: bench.py (kernel M reads/s, e2e M reads/s, roofline fraction) for the in-tree
# library and for each variants/<name>.so built by tools/build_variant.sh
for v in main "$@"; do
  if [ $v = main ]; then unset BSMAP_B200_LIB; else export BSMAP_B200_LIB=variants/$v.so; fi
  python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/b_$v.log 2>&1
  tail -1 gpurun_out/b_$v.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e6,1), round(d['e2e']['value']/1e6,1), round(d['roofline']['frac'],3))"
done
