#!/bin/bash
# deposit modes on one box, interleaved twice: 1 = fp64 CAS.128 pairs, 4 = fixed point (native 32-bit adds), 3 = warp-private
for rep in 1 2; do
  for dep in 1 4 3; do
    for arith in strict tolerance; do
      python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sustained-steps 0 --no-alt-arith --deposit $dep --arith $arith "$@" > gpurun_out/dab_${dep}_${arith}_$rep.json 2> gpurun_out/dab_${dep}_${arith}_$rep.err
      python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/dab_${dep}_${arith}_$rep.json").read().strip().splitlines()[-1])
    print("dep $dep $arith $rep step %.4f irk1 %.4f irk2 %.4f frac %.3f"%(d["ms_per_step"], d["roofline_detail"]["irk1"]["ms_per_launch"], d["roofline"]["ms_per_launch"], d["roofline_detail"]["step"]["frac"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "mode", d["deposit_mode"])
except Exception as e:
    print("dep $dep $arith $rep ERR", e, open("gpurun_out/dab_${dep}_${arith}_$rep.err").read()[-600:])
PY
    done
  done
done
