#!/bin/bash
# sustained (power-capped) step time of the arithmetic / deposit combinations: 500 steps after the 20-step burst
for cfg in "strict 4" "tolerance 4" "tolerance 1" "strict 1"; do
  set -- $cfg
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-alt-arith --sustained-steps 500 --arith $1 --deposit $2 > gpurun_out/sus_$1_$2.json 2> gpurun_out/sus_$1_$2.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sus_$1_$2.json").read().strip().splitlines()[-1])
    print("$1 dep $2 burst %.4f sustained %.4f ms/step"%(d["ms_per_step"], d["sustained"]["ms_per_step"]), d["sustained"]["clocks"]["sm_mhz"], d["sustained"]["clocks"]["reasons"])
except Exception as e:
    print("$1 $2 ERR", e, open("gpurun_out/sus_$1_$2.err").read()[-300:])
PY
done
