"""Configuration sweep behind profiles/r02_config_sweep.md: bench.py at the BASELINE.json-like sizes other than the bench
default, per deposit strategy / arithmetic mode, with the timestep replayed as a CUDA graph and launched kernel by kernel.
Run on a B200 from the repo root:  python tools_py3/config_sweep.py"""
import json
import subprocess
import sys

CFGS = [("C1 default 6.4e6 nx=192", 6.4e6, 192, (0, 4)), ("C2 1e7 nx=256", 1e7, 256, (0, 4)),
        ("C3-like 1e7 nx=4096", 1e7, 4096, (1, 4)), ("C5-like 1.25e8 nx=8192", 1.25e8, 8192, (1, 4))]
for name, n, nx, deps in CFGS:
    for dep in deps:
        for extra in ([], ["--no-graph"], ["--arith", "tolerance"]):
            if extra == ["--no-graph"] and n > 2e7:
                continue
            cmd = [sys.executable, "bench.py", "--steps", "50", "--warmup", "5", "--no-cpu-baseline", "--no-e2e", "--sustained-steps", "0",
                   "--no-launch-timing", "--markers", str(n), "--nx", str(nx), "--deposit", str(dep)] + extra
            r = subprocess.run(cmd, capture_output=True, text=True)
            try:
                d = json.loads(r.stdout.strip().splitlines()[-1])
                rd = d["roofline_detail"]
                print(f"{name:26s} dep={d['deposit_mode']} {' '.join(extra) or 'graph':18s} thr={d['cta_threads']:4d} ctas={d['grid_ctas']:3d} "
                      f"value={d['value']:.3e} ms/step={d['ms_per_step']:.4f} step_frac={rd['step']['frac']:.3f} "
                      f"irk2={d['roofline']['frac']:.3f} irk1={rd['irk1']['frac']:.3f} "
                      f"grid_ms={rd['grid_kernels_ms']['reduce+allreduce+finalize']:.3f}+{rd['grid_kernels_ms']['field_solve']:.3f}", flush=True)
            except Exception as e:
                print(name, dep, extra, "FAILED", r.stderr[-300:], flush=True)
