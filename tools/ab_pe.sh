for v in main "$@"; do
  if [ $v = main ]; then unset BSMAP_B200_LIB; else export BSMAP_B200_LIB=variants/$v.so; fi
  python bench_configs.py --configs cfg3 --steps 2 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['kernel_pairs_per_s']/1e6,2), round(d['e2e_pairs_per_s']/1e6,2), d['paired_fraction'])"
done
