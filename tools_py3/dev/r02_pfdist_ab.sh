for rep in 1 2; do for arith in strict tolerance; do for lib in "" $PWD/scratch/libPFD2.so; do
tag=$(basename "${lib:-tree}" .so)
PIC1DP_B200_LIB=$lib python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-alt-arith --sustained-steps 0 --arith $arith > gpurun_out/pfd_$tag.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/pfd_$tag.json").read().strip().splitlines()[-1])
print("$tag $arith $rep step %.4f irk1 %.4f irk2 %.4f"%(d["ms_per_step"], d["roofline_detail"]["irk1"]["ms_per_launch"], d["roofline"]["ms_per_launch"]), d["deposit_mode"])
PY
done; done; done
