#!/bin/bash
# ncu --set full of the fused output kernel at 2e7 markers
TAG=${1:-r02_prof_diag}
C2="python bench.py --steps 2 --warmup 1 --markers 2e7 --no-cpu-baseline --sustained-steps 0"
ncu --set full --clock-control none --import-source on -k regex:k_diag_fused -s 1 -c 1 -f -o gpurun_out/${TAG} $C2 > gpurun_out/${TAG}_ncu.log 2>&1
python tools_py3/ncu_summary.py gpurun_out/${TAG}.ncu-rep > gpurun_out/${TAG}.md
rm -f gpurun_out/${TAG}.ncu-rep
