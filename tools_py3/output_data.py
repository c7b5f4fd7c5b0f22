"""Python-3 restatement of the reference's output reader (/root/reference/tools/OutputData.py:26-150 together with
the PetscBinaryIO calls it makes), for analysing `pic1dp.out` files.  Analysis / test infrastructure."""
from __future__ import annotations

import os

import numpy as np

from .runinfo import findpeak_energy, growthrate_energy_fit

VEC_FILE_CLASSID = 1211214


class OutputData:
    def __init__(self, datapath):
        path = datapath if os.path.isfile(datapath) else os.path.join(datapath, "pic1dp.out")
        with open(path, "rb") as f:
            ri = lambda n: np.fromfile(f, dtype=">i4", count=n)
            rr = lambda n: np.fromfile(f, dtype=">f8", count=n)
            self.nspecies, self.nmode, self.nx, self.nv, self.nx_pd, self.nv_pd = (int(t) for t in ri(6))
            self.mode = ri(self.nmode)
            self.lx, self.v_max = (float(t) for t in rr(2))
            self.x = np.arange(self.nx + 1.0) / self.nx * self.lx
            self.x_pd = np.arange(self.nx_pd + 1.0) / self.nx_pd * self.lx
            self.v_pd = (np.arange(self.nv_pd + 0.0) / (self.nv_pd - 1) - 0.5) * 2.0 * self.v_max
            self._raw = []
            while True:
                sc = rr(self.nspecies * 3 + 2)
                if sc.size < self.nspecies * 3 + 2:
                    break
                rec = [sc]
                ok = True
                for _ in range(4):
                    hdr = ri(2)
                    if hdr.size < 2 or hdr[0] != VEC_FILE_CLASSID:
                        ok = False
                        break
                    rec.append(rr(int(hdr[1])))
                if not ok:
                    break
                for _ in range(self.nspecies):
                    for _ in range(3):
                        rec.append(rr(self.nx_pd * self.nv_pd))
                    for _ in range(3):
                        rec.append(rr(self.nv_pd))
                if rec[-1].size < self.nv_pd:
                    break
                self._raw.append(rec)
        self.ntime = len(self._raw)

    def get_scalar_t(self):
        out = np.zeros(((self.nspecies + 1) * 3 + 2, self.ntime))
        for it, rec in enumerate(self._raw):
            out[: self.nspecies * 3 + 2, it] = rec[0]
            for s in range(self.nspecies):
                for q in range(3):
                    out[self.nspecies * 3 + 2 + q, it] += rec[0][s * 3 + 2 + q]
        return out

    def get_mode_t(self):
        out = np.zeros((self.nmode * 2, self.ntime))
        for it, rec in enumerate(self._raw):
            out[: self.nmode, it] = rec[1]
            out[self.nmode:, it] = rec[2]
        return out

    def get_field_x(self, itime):
        out = np.zeros((2, self.nx + 1))
        out[0, : self.nx] = self._raw[itime][3]
        out[1, : self.nx] = self._raw[itime][4]
        out[:, self.nx] = out[:, 0]
        return out

    def get_ptcldist_xv(self, itime, ispecies, iptcldist):
        return self._raw[itime][5 + ispecies * 6 + iptcldist].reshape((self.nv_pd, self.nx_pd))

    def get_ptcldist_v(self, itime, ispecies, iptcldist):
        return self._raw[itime][8 + ispecies * 6 + iptcldist]

    def growthrate_energy_fit(self, time1, time2):
        sc = self.get_scalar_t()
        return growthrate_energy_fit(sc[0], sc[1], time1, time2)

    def findpeak_energy(self, time1, time2):
        sc = self.get_scalar_t()
        return findpeak_energy(sc[0], sc[1], time1, time2)
