# tools/ab_bench.sh [variant ...]: bench.py (kernel M units/s, e2e M units/s, roofline fraction) for the in-tree
# library and for each variants/<name>.so built by tools/build_variant.sh; CFG=cfgN selects the workload
for v in main "$@"; do
  if [ $v = main ]; then unset BSMAP_B200_LIB; else export BSMAP_B200_LIB=variants/$v.so; fi
  python bench.py --config ${CFG:-cfg2} --steps 3 --warmup 3 --no-cpu > gpurun_out/b_$v.log 2>&1
  tail -1 gpurun_out/b_$v.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']/1e6,2), round(d['e2e']['value']/1e6,2), round(d['e2e']['ascii']['value']/1e6,2), round(d['roofline']['frac'],3), d['e2e']['records_identical_to_resident_run'])" || tail -5 gpurun_out/b_$v.log
done
