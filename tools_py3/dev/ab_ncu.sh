#!/bin/bash
for lib in "" $(ls scratch/lib*.so 2>/dev/null); do
  tag=$(basename "${lib:-tree}" .so)
  PIC1DP_B200_LIB=${lib:+$PWD/$lib} ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_push -c 4 --csv --log-file gpurun_out/abncu_${tag}.csv python bench.py --steps 2 --warmup 1 --markers 2e7 --no-cpu-baseline --no-e2e "$@" > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/abncu_${tag}.csv")) if len(r)>10]
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ii=h.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((r[ii],r[ki][:40]),{})[r[mi]]=r[vi]
for k,v in d.items(): print("${tag}",k,v)
PY
done
