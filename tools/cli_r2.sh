# tools/cli_r2.sh: whole-process wall clock of the bsmap command line (round 2), one GPU box
python tests/cli_bench.py --reads 20000000 --len 100 --genome-mb 3100 --opts "-s 16 -v 5 -I 4 -S 7" --skip-ref --repeat 2 > gpurun_out/cli_r2_cfg2.json 2> gpurun_out/cli_r2_cfg2.err; tail -c 1500 gpurun_out/cli_r2_cfg2.json
python tests/cli_bench.py --reads 4000000 --len 100 --genome-mb 200 --opts "-s 16 -v 5 -I 4 -S 7" --repeat 3 > gpurun_out/cli_r2_200mb.json 2> gpurun_out/cli_r2_200mb.err; tail -c 1200 gpurun_out/cli_r2_200mb.json
