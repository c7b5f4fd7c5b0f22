#!/bin/bash
# ncu --set full of the fused kernels for a deposit / arithmetic variant at 2e7 markers (summarised on the box)
# usage: r02_prof.sh TAG "bench args"
TAG=$1; shift
C2="python bench.py --steps 2 --warmup 1 --markers 2e7 --no-cpu-baseline --no-e2e --no-graph $*"
$C2 > /dev/null 2> gpurun_out/${TAG}.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_push -s 2 -c 2 -f -o gpurun_out/${TAG} $C2 > gpurun_out/${TAG}_ncu.log 2>&1
python tools_py3/ncu_summary.py gpurun_out/${TAG}.ncu-rep > gpurun_out/${TAG}.md
rm -f gpurun_out/${TAG}.ncu-rep
