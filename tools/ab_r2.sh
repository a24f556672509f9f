bash tools/ab_bench.sh se12
CFG=cfg4 bash tools/ab_bench.sh rrbs12 rrbs16
CFG=cfg3 bash tools/ab_bench.sh
