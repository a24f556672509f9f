# tools/ncu_kernel.sh TAG CONFIG KERNEL_REGEX UNITS [SKIP]: one `ncu --set full` capture of a mapping / index kernel while bench.py runs CONFIG
# with UNITS reads (pairs) per launch -> gpurun_out/prof_TAG.ncu-rep   (never a bench value: ncu serialises and replays the kernel)
tag=$1; cfg=$2; k=$3; units=$4; skip=${5:-1}
ncu --set full --import-source on --clock-control none -k regex:$k -s $skip -c 1 -f -o gpurun_out/prof_$tag \
    python bench.py --config $cfg --steps 1 --warmup 1 --no-cpu --reads $units > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log | cut -c1-200
