"""Configuration sweep behind profiles/r01_config_sweep.md: bench.py at every BASELINE.json-like size x deposit strategy.
Run on a B200 from the repo root:  python tools_py3/config_sweep.py"""
import sys, json, subprocess
cfgs=[("C1 default", 6.4e6, 192), ("C2", 1e7, 256), ("C3-like nx=4096", 1e7, 4096), ("C4", 1e8, 1024), ("C5-like nx=8192 1.25e8", 1.25e8, 8192), ("big 4e8 nx=1024", 4e8, 1024)]
for name,n,nx in cfgs:
    for dep in (1,3,2):
        r=subprocess.run([sys.executable,"bench.py","--steps","20","--warmup","3","--no-cpu-baseline","--no-e2e","--markers",str(n),"--nx",str(nx),"--deposit",str(dep)],capture_output=True,text=True)
        try:
            d=json.loads(r.stdout.strip().splitlines()[-1])
            print(f"{name:28s} dep={d['deposit_mode']} thr={d['cta_threads']:4d} ctas={d['grid_ctas']:3d} smem={d['smem_bytes']:6d} value={d['value']:.3e} ms/step={d['ms_per_step']:.3f} step_frac={d['roofline_detail']['step']['frac']:.3f} irk2={d['roofline']['frac']:.3f} irk1={d['roofline_detail']['irk1']['frac']:.3f} grid_ms={d['roofline_detail']['grid_kernels_ms']}", flush=True)
        except Exception as e:
            print(name, dep, "FAILED", r.stderr[-300:], flush=True)
