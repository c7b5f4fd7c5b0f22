#!/bin/bash
# A/B of kernel variants on one box: the in-tree library and scratch/lib*.so, interleaved twice
for rep in 1 2; do
  for lib in "" $(ls scratch/lib*.so 2>/dev/null); do
    tag=$(basename "${lib:-tree}" .so)
    PIC1DP_B200_LIB=${lib:+$PWD/$lib} python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ab_${tag}_$rep.json 2> gpurun_out/ab_${tag}_$rep.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${tag}_$rep.json").read().strip().splitlines()[-1])
    print("${tag}", $rep, "step %.4f irk1 %.4f irk2 %.4f"%(d["ms_per_step"], d["roofline_detail"]["irk1"]["ms_per_launch"], d["roofline"]["ms_per_launch"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "dep", d["deposit_mode"])
except Exception as e:
    print("${tag}", $rep, "ERR", e, open("gpurun_out/ab_${tag}_$rep.err").read()[-600:])
PY
  done
done
