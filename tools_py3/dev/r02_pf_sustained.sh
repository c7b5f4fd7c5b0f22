#!/bin/bash
# L2 prefetch of the next tile step in the irk=1 kernel (PIC1DP_PF_MASK bit 0): burst and power-capped step time
for rep in 1 2; do
  for lib in "" $PWD/scratch/libPF12.so; do
    tag=$(basename "${lib:-tree}" .so)_$rep
    PIC1DP_B200_LIB=$lib python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-alt-arith --sustained-steps 500 > gpurun_out/pfs_$tag.json 2> gpurun_out/pfs_$tag.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/pfs_$tag.json").read().strip().splitlines()[-1])
    print("$tag burst %.4f (irk1 %.4f irk2 %.4f) sustained %.4f ms/step"%(d["ms_per_step"], d["roofline_detail"]["irk1"]["ms_per_launch"], d["roofline"]["ms_per_launch"], d["sustained"]["ms_per_step"]), d["sustained"]["clocks"]["sm_mhz"], d["sustained"]["clocks"]["reasons"])
except Exception as e:
    print("$tag ERR", e, open("gpurun_out/pfs_$tag.err").read()[-300:])
PY
  done
done
