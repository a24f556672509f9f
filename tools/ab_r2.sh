CFG=cfg3 bash tools/ab_bench.sh pe_ku4 2>&1 | grep -v "^$"
CFG=cfg2 bash tools/ab_bench.sh se_ku4 2>&1 | grep -v "^$"
bash tools/ncu_kernel.sh r2d_pe cfg3 bsx_map_pe 2000000
