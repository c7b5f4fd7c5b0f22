"""Distribution data files of the reference's tools/ptcldist.py, and a marker sampler that reads them.

The reference tool is an EXPORTER (/root/reference/tools/ptcldist.py:60-73): it reads `pic1dp.out` and writes the
binned distribution of one output time as text -- `ptcldist_xv_<t>_<s>_<d>.dat` (np.savetxt of the nv_pd x (nx_pd+1)
array from OutputData.get_ptcldist_xv, periodic column appended), `x_*.dat` (nx_pd+1 grid points) and `v_*.dat`
(nv_pd grid points).  The reference has no particle-file input; BASELINE.json configs[2]
("markers from a ptcldist.py file") is therefore served by the builder-defined sampler below: markers are drawn
uniformly in (x, v) like the reference's uniform-v loading (src/pic1dp_particle.F90:179-181, 222-223) and carry
p = f(x, v) / g with f interpolated bilinearly from the file.
"""
from __future__ import annotations

import os

import numpy as np


def export_xv(od, itime: int, ispecies: int, idist: int, outdir: str = ".", tag: str = ""):
    """Write ptcldist_xv / x / v files exactly as tools/ptcldist.py does for `-xv 0 -t itime -s ispecies -d idist`."""
    ext = "_%d_%d_%d" % (itime, ispecies, idist) + (("_" + tag) if tag else "") + ".dat"
    pd = od.get_ptcldist_xv(itime, ispecies, idist)                 # (nv_pd, nx_pd)
    pd = np.concatenate([pd, pd[:, :1]], axis=1)                    # periodic boundary column (OutputData.py:123-131)
    paths = {k: os.path.join(outdir, k + ext) for k in ("ptcldist_xv", "x", "v")}
    np.savetxt(paths["ptcldist_xv"], pd)
    np.savetxt(paths["x"], od.x_pd)
    np.savetxt(paths["v"], od.v_pd)
    return paths


def load_xv(paths):
    pd = np.loadtxt(paths["ptcldist_xv"])
    x = np.loadtxt(paths["x"])
    v = np.loadtxt(paths["v"])
    assert pd.shape == (v.size, x.size), (pd.shape, v.size, x.size)
    return pd, x, v


def sample_markers(pd, xg, vg, n: int, seed: int = 0):
    """n markers uniform in x in [0, lx) and v in [v_min, v_max] with p = lx (v_max - v_min) / n * f(x, v), f bilinear
    in the file's grid.  Returns x, v, p (w is the caller's perturbation)."""
    rng = np.random.default_rng(seed)
    lx, vmin, vmax = xg[-1], vg[0], vg[-1]
    x = rng.random(n) * lx
    v = vmin + rng.random(n) * (vmax - vmin)
    sx = x / lx * (xg.size - 1)
    ix = np.minimum(np.floor(sx).astype(np.int64), xg.size - 2)
    fx = sx - ix
    sv = (v - vmin) / (vmax - vmin) * (vg.size - 1)
    iv = np.minimum(np.floor(sv).astype(np.int64), vg.size - 2)
    fv = sv - iv
    f = (pd[iv, ix] * (1 - fx) * (1 - fv) + pd[iv, ix + 1] * fx * (1 - fv) +
         pd[iv + 1, ix] * (1 - fx) * fv + pd[iv + 1, ix + 1] * fx * fv)
    p = lx * (vmax - vmin) / n * f
    return x, v, p
