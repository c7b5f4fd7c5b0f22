#!/bin/bash
# the fixed-point deposit (native adds) with serial / overlapped deposits of a thread's two markers, and the fp64 CAS deposit
one() { # tag lib dep arith
  PIC1DP_B200_LIB=$2 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sustained-steps 0 --no-alt-arith --deposit $3 --arith $4 > gpurun_out/dab2_$1.json 2> gpurun_out/dab2_$1.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/dab2_$1.json").read().strip().splitlines()[-1])
    print("$1 step %.4f irk1 %.4f irk2 %.4f frac %.3f"%(d["ms_per_step"], d["roofline_detail"]["irk1"]["ms_per_launch"], d["roofline"]["ms_per_launch"], d["roofline_detail"]["step"]["frac"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "mode", d["deposit_mode"])
except Exception as e:
    print("$1 ERR", e, open("gpurun_out/dab2_$1.err").read()[-600:])
PY
}
for rep in 1 2; do
  for arith in strict tolerance; do
    one cas_${arith}_$rep "" 1 $arith
    one fixed_${arith}_$rep "" 4 $arith
    one fixedadd2_${arith}_$rep $PWD/scratch/libADD2F.so 4 $arith
  done
done
