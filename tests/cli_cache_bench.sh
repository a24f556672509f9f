# tests/cli_cache_bench.sh: config 2 at full size through the command line, without and with the packed reference cache
python tests/cli_bench.py --skip-ref --repeat 1 --reads 20000000 --len 100 --genome-mb 3100 --opts "-s 16 -v 5 -I 4 -S 7" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no cache', d['ours_first_run_seconds'], d['ours_all_runs_seconds'], d['ours_stages'])"
mkdir -p /tmp/bsxcache
BSX_REF_CACHE=/tmp/bsxcache python tests/cli_bench.py --skip-ref --repeat 2 --reads 20000000 --len 100 --genome-mb 3100 --opts "-s 16 -v 5 -I 4 -S 7" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cache: first (writes it)', d['ours_first_run_seconds'], 'then', d['ours_all_runs_seconds'], d['ours_stages']); print(d['ours_md5'])"
ls -la /tmp/bsxcache
