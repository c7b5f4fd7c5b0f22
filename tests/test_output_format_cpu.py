"""pic1dp.out layout: a file written with the documented layout (src/pic1dp_output.F90:74-92, :117-187, :457-474)
is parsed by the py3 reader restatement; growth-rate fit on a synthetic exponential."""
import os

import numpy as np

from tools_py3.output_data import OutputData, VEC_FILE_CLASSID
from tools_py3.runinfo import findpeak_energy, growthrate_energy_fit, intfdt


def _write(path, nt=30, nx=8, nspecies=1, nmode=1, nxo=4, nvo=4):
    with open(path, "wb") as f:
        np.asarray([nspecies, nmode, nx, 128, nxo, nvo, 1], dtype=">i4").tofile(f)
        np.asarray([17.45, 8.0], dtype=">f8").tofile(f)
        for it in range(nt):
            t = 0.5 * it
            np.asarray([t, 1e-9 * np.exp(0.2 * t), 1.0, 2.0, 3.0], dtype=">f8").tofile(f)
            for n in (nmode, nmode, nx, nx):
                np.asarray([VEC_FILE_CLASSID, n], dtype=">i4").tofile(f)
                np.asarray(np.full(n, t), dtype=">f8").tofile(f)
            for n in (nxo * nvo,) * 3 + (nvo,) * 3:
                np.asarray(np.full(n, it), dtype=">f8").tofile(f)


def test_reader_roundtrip(tmp_path):
    p = tmp_path / "pic1dp.out"
    _write(p)
    od = OutputData(str(p))
    assert od.ntime == 30 and od.nx == 8 and od.lx == 17.45
    sc = od.get_scalar_t()
    assert sc.shape == (8, 30) and sc[5, 3] == 1.0 and sc[7, 3] == 3.0
    assert od.get_field_x(4)[0, 0] == 2.0 and od.get_field_x(4)[0, 8] == 2.0
    assert od.get_ptcldist_xv(7, 0, 2).shape == (4, 4) and od.get_ptcldist_v(7, 0, 1)[0] == 7.0
    assert abs(od.growthrate_energy_fit(2.0, 10.0) - 0.2) < 1e-12


def test_fit_helpers():
    t = np.linspace(0, 10, 101)
    e = 3e-9 * np.exp(0.1677 * t)
    assert abs(growthrate_energy_fit(t, e, 2.0, 8.0) - 0.1677) < 1e-12
    assert findpeak_energy(t, e, 0.0, 10.0)[0] == t[-2] or findpeak_energy(t, e, 0.0, 10.1)[0] == 10.0
    assert abs(intfdt(t, np.ones_like(t)) - 10.0) < 1e-12


def test_ptcldist_files_roundtrip_and_sampler(tmp_path):
    """tools_py3/ptcldist.py: same file layout as the reference exporter (nv_pd rows x (nx_pd + 1) columns with the
    periodic column, x and v grids), and the marker sampler integrates the binned f back to the density."""
    from tools_py3 import ptcldist
    p = tmp_path / "pic1dp.out"
    _write(p, nt=2, nxo=8, nvo=6)
    od = OutputData(str(p))
    paths = ptcldist.export_xv(od, 1, 0, 1, outdir=str(tmp_path), tag="run")
    assert os.path.basename(paths["ptcldist_xv"]) == "ptcldist_xv_1_0_1_run.dat"
    pd, xg, vg = ptcldist.load_xv(paths)
    assert pd.shape == (6, 9) and np.all(pd[:, -1] == pd[:, 0]) and xg.size == 9 and vg.size == 6
    # a Maxwellian on a 64 x 65 grid, sampled back
    vg = (np.arange(64) / 63.0 - 0.5) * 16.0
    xg = np.arange(65) / 64.0 * 12.0
    pd = np.repeat((np.exp(-vg * vg / 2) / np.sqrt(2 * np.pi))[:, None], 65, axis=1)
    x, v, pw = ptcldist.sample_markers(pd, xg, vg, 200000, seed=1)
    assert 0.0 <= x.min() and x.max() < 12.0 and abs(v).max() <= 8.0
    assert abs(np.sum(pw) / 12.0 - 1.0) < 0.02 and abs(np.sum(pw * v * v) / np.sum(pw) - 1.0) < 0.05
