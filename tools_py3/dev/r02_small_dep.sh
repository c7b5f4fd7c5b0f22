#!/bin/bash
# small grids: warp-private (AUTO for nx <= 256) against the fixed-point deposit with native adds and the CAS deposit
for rep in 1 2; do
  for cfg in "c0 --markers 6.4e6 --nx 192" "c1 --markers 1e7 --nx 256" "c5 --markers 1e7 --nx 512" "c2 --markers 1e7 --nx 4096"; do
    set -- $cfg; name=$1; shift
    for dep in 3 4 1; do
      python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-e2e --sustained-steps 0 --no-launch-timing --no-alt-arith --deposit $dep "$@" > gpurun_out/sd_${name}_$dep.json 2> gpurun_out/sd_${name}_$dep.err
      python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sd_${name}_$dep.json").read().strip().splitlines()[-1])
    print("$name dep $dep rep $rep ms/step %.4f frac %.3f mode"%(d["ms_per_step"], d["roofline_detail"]["step"]["frac"]), d["deposit_mode"], d["cta_threads"])
except Exception as e:
    print("$name dep $dep ERR", e, open("gpurun_out/sd_${name}_$dep.err").read()[-300:])
PY
    done
  done
done
