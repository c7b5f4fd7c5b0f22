"""Growth-rate and peak diagnostics, restated for Python 3 from the reference's Python-2 tools.

growthrate_energy_fit follows /root/reference/tools/OutputData.py:153-170 (least-squares slope of ln(energy) on
[searchsorted(t, t1) - 1, searchsorted(t, t2))); tools/runinfo.py:116 halves it to get the amplitude growth rate.
findpeak_energy follows tools/OutputData.py:172-180; intfdt follows tools/runinfo.py:30-37.
"""
from __future__ import annotations

import numpy as np


def growthrate_energy_fit(t, energy, time1, time2):
    t = np.asarray(t, dtype=np.float64)
    energy = np.asarray(energy, dtype=np.float64)
    itime1 = max(int(np.searchsorted(t, time1)) - 1, 0)
    itime2 = int(np.searchsorted(t, time2))
    tt = t[itime1:itime2]
    ln = np.log(energy[itime1:itime2])
    n = itime2 - itime1
    return (n * np.sum(tt * ln) - np.sum(tt) * np.sum(ln)) / (n * np.sum(tt * tt) - np.sum(tt) * np.sum(tt))


def findpeak_energy(t, energy, time1, time2):
    t = np.asarray(t)
    energy = np.asarray(energy)
    itime1 = max(int(np.searchsorted(t, time1)) - 1, 0)
    itime2 = int(np.searchsorted(t, time2))
    k = int(np.argmax(energy[itime1:itime2]))
    return float(t[itime1 + k]), float(energy[itime1 + k])


def intfdt(t, f):
    t = np.asarray(t)
    f = np.asarray(f)
    nt = len(t)
    integral = (t[1] - t[0]) * f[0] + (t[nt - 1] - t[nt - 2]) * f[nt - 1]
    integral += np.sum((t[2:] - t[:-2]) * f[1:-1])
    return integral / 2.0
