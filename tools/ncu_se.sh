# tools/ncu_se.sh TAG [kernel regex]: one `ncu --set full` capture of the SE mapping kernel (4 M reads, config-2 genome)
# -> gpurun_out/prof_TAG.ncu-rep   (never a bench value: ncu serialises and replays the kernel)
tag=$1; k=${2:-bsx_map_se}
ncu --set full --import-source on --clock-control none -k regex:$k -s 1 -c 1 -f -o gpurun_out/prof_$tag \
    python bench.py --steps 1 --warmup 1 --no-cpu --reads 4000000 > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log | cut -c1-300
