#!/bin/bash
# Round-end measurement recipe on a B200 box (run from the repo root through gpurun):
#   plain bench (exit 0) -> ncu launch list of the same command -> one `ncu --set full` capture per dominant kernel.
# Outputs go to gpurun_out/ (scratch); tools_py3/launch_list.py and tools_py3/ncu_summary.py turn them into profiles/*.md.
set -x
TAG=${1:-r01}
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --sustained-steps 0 --no-alt-arith --no-graph"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
for dep in 4 1; do
  C2="python bench.py --steps 2 --warmup 1 --markers 2e7 --no-cpu-baseline --no-e2e --deposit $dep"
  $C2 > /dev/null 2>&1 || exit 1
  ncu --set full --clock-control none --import-source on -k regex:k_push -s 2 -c 2 -f -o gpurun_out/${TAG}_prof_push_dep$dep $C2 > gpurun_out/${TAG}_ncu_full_dep$dep.log 2>&1
  # summarise on the box: gpurun brings back at most 64 MiB, one .ncu-rep is ~40 MB
  python tools_py3/ncu_summary.py gpurun_out/${TAG}_prof_push_dep$dep.ncu-rep > gpurun_out/${TAG}_ncu_push_dep$dep.md
  rm -f gpurun_out/${TAG}_prof_push_dep$dep.ncu-rep
done
ls -la gpurun_out/${TAG}_*
