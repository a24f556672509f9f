CFG=cfg2 bash tools/ab_bench.sh se_pf1 se_pf2 2>&1 | grep -v "^$"
grep -o '"value_ascii_resident": [0-9.]*' gpurun_out/b_main.log
CFG=cfg3 bash tools/ab_bench.sh pe_pf2 2>&1 | grep -v "^$"
bash tools/ncu_se.sh r2e
