# tools/ab_cfg.sh CFG [variant ...]: bench_configs.py rate of one config for the in-tree library and variants/<name>.so
cfg=$1; shift
for v in main "$@"; do
  if [ $v = main ]; then unset BSMAP_B200_LIB; else export BSMAP_B200_LIB=variants/$v.so; fi
  python bench_configs.py --configs $cfg --steps 2 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', {k: (round(x,3) if isinstance(x,float) else x) for k,x in d.items() if 'per_s' in k})"
done
