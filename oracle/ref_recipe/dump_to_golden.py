#!/usr/bin/env python3
"""dump_to_golden.py DUMPDIR OUT.npz -- converts the refdump_rank*.bin files written by the reference built with
oracle/ref_recipe/pic1dp_refdump.F90 into one compressed fixture: per record (itime, irk) the per-rank marker arrays
and the assembled rho, E, mode_re, mode_im.  Layout: see the header of pic1dp_refdump.F90."""
import glob
import os
import sys

import numpy as np


def read_rank(path):
    recs = []
    with open(path, "rb") as f:
        buf = f.read()
    o = 0

    def take(dt, n):
        nonlocal o
        a = np.frombuffer(buf, dtype=dt, count=n, offset=o)
        o += a.nbytes
        return a
    while o < len(buf):
        magic, itime, irk, npe, mype, nx, nmode, nsp = (int(t) for t in take("<i4", 8))
        assert magic == 20140512, "bad magic (endianness / layout?)"
        r = dict(itime=itime, irk=irk, npe=npe, mype=mype, nx=nx, nmode=nmode, species=[])
        for _ in range(nsp):
            n = int(take("<i8", 1)[0])
            r["species"].append({k: take("<f8", n).copy() for k in ("x", "v", "p", "w")})
        lo, hi = (int(t) for t in take("<i4", 2))
        r["ix"] = (lo, hi)
        r["rho"], r["E"] = take("<f8", hi - lo).copy(), take("<f8", hi - lo).copy()
        lo, hi = (int(t) for t in take("<i4", 2))
        r["im"] = (lo, hi)
        r["mode_re"], r["mode_im"] = take("<f8", hi - lo).copy(), take("<f8", hi - lo).copy()
        recs.append(r)
    return recs


def main():
    d, out = sys.argv[1], sys.argv[2]
    ranks = [read_rank(p) for p in sorted(glob.glob(os.path.join(d, "refdump_rank*.bin")))]
    assert ranks and all(len(r) == len(ranks[0]) for r in ranks)
    npe, nx, nmode = ranks[0][0]["npe"], ranks[0][0]["nx"], ranks[0][0]["nmode"]
    assert npe == len(ranks)
    z = dict(npe=npe, nx=nx, nmode=nmode, nrec=len(ranks[0]), nspecies=len(ranks[0][0]["species"]))
    for k, recs in enumerate(zip(*ranks)):
        z[f"r{k}_itime"], z[f"r{k}_irk"] = recs[0]["itime"], recs[0]["irk"]
        rho, E, mre, mim = np.zeros(nx), np.zeros(nx), np.zeros(nmode), np.zeros(nmode)
        for r in recs:
            rho[r["ix"][0]:r["ix"][1]], E[r["ix"][0]:r["ix"][1]] = r["rho"], r["E"]
            mre[r["im"][0]:r["im"][1]], mim[r["im"][0]:r["im"][1]] = r["mode_re"], r["mode_im"]
            for s, sp in enumerate(r["species"]):
                for q in ("x", "v", "p", "w"):
                    z[f"r{k}_rank{r['mype']}_s{s}_{q}"] = sp[q]
        z[f"r{k}_rho"], z[f"r{k}_E"], z[f"r{k}_mode_re"], z[f"r{k}_mode_im"] = rho, E, mre, mim
    np.savez_compressed(out, **z)
    print("records:", len(ranks[0]), "ranks:", npe, "->", out)


if __name__ == "__main__":
    main()
