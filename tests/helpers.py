"""Shared test helpers: synthetic markers and the oracle-side replay of the reference time loop."""
from __future__ import annotations

import numpy as np

from oracle import oracle as O
import pic1dp_b200 as P

COMMON = ("nx", "nmode", "modes", "lx", "dt", "nspecies", "charge", "mass", "temperature", "temperature2",
          "density", "v0", "iptcldist", "deltaf", "linear", "iptclshape")


def make_params(**over):
    """Matching (oracle params, GPU params) pair.  GPU-only keys: capacity, device, deposit_mode, field_mode, fuse,
    rank, nranks."""
    gpu_only = {k: over.pop(k) for k in list(over) if k in ("capacity", "device", "deposit_mode", "field_mode",
                                                            "fuse", "rank", "nranks", "load_path", "arith_mode",
                                                            "no_step_graph")}
    op = O.default_params(**over)
    loader_only = ("init_nmode", "init_mode", "init_mode_cos", "init_mode_sin", "v_max", "imarker")
    gp = P.default_params(**{k: v for k, v in over.items() if k not in loader_only}, **gpu_only)
    return op, gp


def synth_markers(op, n, seed=0, isp=0, spread=0.0):
    """Markers distributed like particle_load's default branch (uniform x, uniform v, p = f0/g + w,
    w = 1e-5 sin(kx) p), from numpy's PCG64 with a fixed seed.  spread > 0 pushes a fraction of x outside [0, lx)
    to exercise the wrap."""
    rng = np.random.default_rng(seed)
    lx, vmax = op.lx, op.v_max
    x = rng.random(n) * lx
    if spread > 0:
        x = x + (rng.random(n) - 0.5) * 2.0 * spread * lx
    v = (rng.random(n) - 0.5) * 2.0 * vmax
    T, T2, m = op.temperature[isp], op.temperature2[isp], op.mass[isp]
    dens, v0 = op.density[isp], op.v0[isp]
    if op.iptcldist == 3:
        f0 = dens * np.exp(-v * v / (2 * T / m)) / np.sqrt(2 * np.pi * T / m) + \
            (1 - dens) * np.exp(-(v - v0) ** 2 / (2 * T2 / m)) / np.sqrt(2 * np.pi * T2 / m)
    elif op.iptcldist == 2:
        f0 = dens * (np.exp(-(v + v0) ** 2 / (2 * T / m)) + np.exp(-(v - v0) ** 2 / (2 * T / m))) / \
            np.sqrt(8 * np.pi * T / m)
    elif op.iptcldist == 1:
        f0 = dens * v * v * np.exp(-v * v / 2) / np.sqrt(2 * np.pi)
    else:
        f0 = dens * np.exp(-(v - v0) ** 2 / (2 * T / m)) / np.sqrt(2 * np.pi * T / m)
    p = lx * 2 * vmax / max(n, 1) * f0
    w = 1e-5 * np.sin(2 * np.pi / lx * x) * p
    if not op.linear:
        p = p + w
    return dict(x=x, v=v, p=p, w=w)


def copy_state(st):
    return {k: a.copy() for k, a in st.items()}


class OracleRun:
    """Replays src/pic1dp.F90:63-90 with the oracle's restated subroutines, one emulated rank per entry of
    `ranks` (list[species][rank] of marker dicts)."""

    def __init__(self, op, states):
        self.op = op
        self.o = O.Oracle(op)
        self.st = states
        for sp in states:
            for st in sp:
                n = st["x"].size
                st["xb"], st["vb"], st["wb"] = np.zeros(n), np.zeros(n), np.zeros(n)
        self.rho = np.zeros(op.nx)
        self.E = np.zeros(op.nx)
        self.mode_re = np.zeros(op.nmode)
        self.mode_im = np.zeros(op.nmode)
        self.noob = 0

    def collect_charge(self):
        key = "w" if self.op.deltaf == 1 else "p"
        xs = [[st["x"] for st in sp] for sp in self.st]
        ws = [[st[key] for st in sp] for sp in self.st]
        self.rho, n = self.o.collect_charge(xs, ws)
        self.noob += n

    def solve_field(self):
        self.E, self.mode_re, self.mode_im = self.o.field_solve(self.rho)

    def push(self, irk):
        for s, sp in enumerate(self.st):
            for st in sp:
                self.o.push_species(s, irk, st["x"], st["v"], st["p"], st["w"], st["xb"], st["vb"], st["wb"], self.E)

    def init_field(self):
        self.collect_charge()
        self.solve_field()

    def step(self):
        for irk in (1, 2):
            self.push(irk)
            self.collect_charge()
            self.solve_field()


def rel_err(a, b, scale=None):
    a, b = np.asarray(a), np.asarray(b)
    if a.size == 0:
        return 0.0
    s = np.max(np.abs(b)) if scale is None else scale
    if s == 0:
        s = 1.0
    return float(np.max(np.abs(a - b)) / s)


class ChunkedOracleRun(OracleRun):
    """OracleRun over one species whose markers are split into `nchunks` contiguous blocks -- the reference's own MPI
    layout (PETSC_DECIDE blocks, per-rank private grids summed in rank order) -- with the per-block push running on a
    thread pool (the C restatement releases the GIL), so that the stated configuration sizes (1e7 .. 1e8 markers)
    finish in seconds on the GPU box's host cores."""

    def __init__(self, op, st, nchunks):
        import concurrent.futures as cf
        n = st["x"].size
        bounds = [O.petsc_decide(n, nchunks, r) for r in range(nchunks)]
        self.full = st          # chunk arrays are views into these
        chunks = [{k: st[k][lo:hi] for k in ("x", "v", "p", "w")} for lo, hi in bounds]
        self.backup = {k: np.zeros(n) for k in ("xb", "vb", "wb")}
        super().__init__(op, [chunks])
        for c, (lo, hi) in zip(chunks, bounds):
            for k in ("xb", "vb", "wb"):
                c[k] = self.backup[k][lo:hi]
        self.pool = cf.ThreadPoolExecutor(max_workers=nchunks)

    def push(self, irk):
        E = np.ascontiguousarray(self.E)
        futs = [self.pool.submit(self.o.push_species, 0, irk, c["x"], c["v"], c["p"], c["w"], c["xb"], c["vb"], c["wb"], E)
                for c in self.st[0]]
        for f in futs:
            f.result()

    def cell_index(self):
        """ix of every marker (src/pic1dp_interaction.F90:106-107) from the current (wrapped) x."""
        n = self.full["x"].size
        ix = np.zeros(n, dtype=np.int32)
        pos = 0
        futs = []
        for c in self.st[0]:
            m = c["x"].size
            futs.append(self.pool.submit(lambda xs, out: out.__setitem__(slice(None), self.o.shape(xs.copy())[0]),
                                         c["x"], ix[pos:pos + m]))
            pos += m
        for f in futs:
            f.result()
        return ix
