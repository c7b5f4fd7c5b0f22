#!/bin/bash
# register-staged v of the next tile step (PIC1DP_PV_MASK bit 0 = irk1, bit 1 = irk2 of the atomic deposits) with the
# fixed-point deposit, both arithmetic modes; small grids in tolerance arithmetic (warp-private vs fixed point)
one() { # tag lib args...
  local tag=$1 lib=$2; shift; shift
  PIC1DP_B200_LIB=$lib python bench.py --warmup 3 --no-cpu-baseline --no-e2e --sustained-steps 0 --no-alt-arith "$@" > gpurun_out/pv_$tag.json 2> gpurun_out/pv_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/pv_$tag.json").read().strip().splitlines()[-1])
    rd=d.get("roofline_detail",{})
    print("$tag step %.4f frac %.3f"%(d["ms_per_step"], rd["step"]["frac"]), "irk1 %.4f irk2 %.4f"%(rd["irk1"]["ms_per_launch"], d["roofline"]["ms_per_launch"]) if d["roofline"].get("ms_per_launch") else "", "mode", d["deposit_mode"])
except Exception as e:
    print("$tag ERR", e, open("gpurun_out/pv_$tag.err").read()[-300:])
PY
}
for rep in 1 2; do
  for arith in strict tolerance; do
    one tree_${arith}_$rep "" --steps 20 --deposit 4 --arith $arith
    one pv5_${arith}_$rep $PWD/scratch/libPV5.so --steps 20 --deposit 4 --arith $arith
    one pv7_${arith}_$rep $PWD/scratch/libPV7.so --steps 20 --deposit 4 --arith $arith
  done
  for dep in 3 4; do
    one c0tol_dep${dep}_$rep "" --steps 200 --no-launch-timing --markers 6.4e6 --nx 192 --deposit $dep --arith tolerance
    one c1tol_dep${dep}_$rep "" --steps 200 --no-launch-timing --markers 1e7 --nx 256 --deposit $dep --arith tolerance
  done
done
