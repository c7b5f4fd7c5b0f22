"""ctypes loader for the CPU restatement (oracle/libpic1dp_oracle.so).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package (pic1dp_b200/).
Parity status: "parity unpinned" for the hot path (see pic1dp_oracle.h); the RNG part is pinned by the
reference's known-answer vectors (/root/reference/src/multirand.F90:396-425).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpic1dp_oracle.so")

MAX_SPECIES = 4
MAX_MODES = 64


class OrcParams(C.Structure):
    _fields_ = [
        ("nx", C.c_int32),
        ("nmode", C.c_int32),
        ("modes", C.c_int32 * MAX_MODES),
        ("lx", C.c_double),
        ("dt", C.c_double),
        ("nspecies", C.c_int32),
        ("charge", C.c_double * MAX_SPECIES),
        ("mass", C.c_double * MAX_SPECIES),
        ("temperature", C.c_double * MAX_SPECIES),
        ("temperature2", C.c_double * MAX_SPECIES),
        ("density", C.c_double * MAX_SPECIES),
        ("v0", C.c_double * MAX_SPECIES),
        ("iptcldist", C.c_int32),
        ("deltaf", C.c_int32),
        ("linear", C.c_int32),
        ("iptclshape", C.c_int32),
        ("v_max", C.c_double),
        ("imarker", C.c_int32),
        ("init_nmode", C.c_int32),
        ("init_mode", C.c_int32 * MAX_MODES),
        ("init_mode_cos", C.c_double * MAX_MODES),
        ("init_mode_sin", C.c_double * MAX_MODES),
    ]


class OrcRankState(C.Structure):
    _fields_ = [("np", C.c_int64)] + [(n, C.POINTER(C.c_double)) for n in ("x", "v", "p", "w", "xb", "vb", "wb")]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc -O3 -ffp-contract=off)."""
    import hashlib
    srcs = [os.path.join(_HERE, f) for f in ("pic1dp_oracle.c", "multirand_oracle.c", "optimize_oracle.c", "pic1dp_oracle.h", "Makefile")]
    hh = hashlib.sha256(b"".join(open(s, "rb").read() for s in srcs)).hexdigest()  # mtimes do not survive gpurun
    hpath = _LIB_PATH + ".hash"
    stale = (not os.path.exists(_LIB_PATH)) or (not os.path.exists(hpath)) or open(hpath).read().strip() != hh
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
        with open(hpath, "w") as f:
            f.write(hh)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp = C.POINTER(C.c_double)
        pp = C.POINTER(OrcParams)
        L.orc_params_default.argtypes = [pp]
        L.orc_field_init.argtypes = [pp, dp, dp, dp]
        L.orc_field_solve.argtypes = [pp, dp, dp, dp, dp, dp, dp, dp]
        L.orc_shape.argtypes = [pp, C.c_int64, dp, C.POINTER(C.c_int32), dp, dp, C.c_int]
        L.orc_deposit_species.argtypes = [pp, C.c_int64, dp, dp, dp]
        L.orc_deposit_species.restype = C.c_int64
        L.orc_collect_charge.argtypes = [pp, C.c_int, C.POINTER(C.c_int64), C.POINTER(dp), C.POINTER(dp), dp]
        L.orc_collect_charge.restype = C.c_int64
        L.orc_push_species.argtypes = [pp, C.c_int, C.c_int, C.c_int64, dp, dp, dp, dp, dp, dp, dp, dp]
        L.orc_dlnf0.argtypes = [pp, C.c_int, C.c_double]
        L.orc_dlnf0.restype = C.c_double
        L.orc_field_energy.argtypes = [pp, dp]
        L.orc_field_energy.restype = C.c_double
        L.orc_particle_load.argtypes = [pp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, dp, dp, dp, dp]
        L.orc_output_field.argtypes = [pp, C.c_int, C.POINTER(C.c_int64), C.POINTER(dp), C.POINTER(dp), C.POINTER(dp), dp, dp]
        L.orc_output_ptcldist.argtypes = [pp, C.c_int, C.c_int, C.POINTER(C.c_int64)] + [C.POINTER(dp)] * 4 + \
            [C.c_int, C.c_int, C.c_double] + [dp] * 6
        L.orc_petsc_decide.argtypes = [C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_run.argtypes = [pp, C.c_int, C.POINTER(OrcRankState), C.c_int, dp, dp, dp, dp, dp, C.c_int]
        L.orc_run.restype = C.c_double
        L.orc_multirand_new.restype = C.c_void_p
        L.orc_multirand_free.argtypes = [C.c_void_p]
        L.orc_multirand_seed_default.argtypes = [C.c_void_p, C.c_int]
        L.orc_multirand_init_const.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_multirand_int64.argtypes = [C.c_void_p]
        L.orc_multirand_int64.restype = C.c_int64
        L.orc_multirand_get_seeds4.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.orc_multirand_skip.argtypes = [C.c_void_p, C.c_int64]
        L.orc_multirand_real64.argtypes = [C.c_void_p]
        L.orc_multirand_real64.restype = C.c_double
        L.orc_multirand_real_array.argtypes = [C.c_void_p, dp, C.c_int64]
        L.orc_multirand_gaussian_array.argtypes = [C.c_void_p, dp, C.c_int64]
        L.orc_dist_pertb_abs_v.argtypes = [C.c_int, C.POINTER(C.c_int64), C.POINTER(dp), C.POINTER(dp), C.c_int, C.c_double, dp]
        L.orc_particle_merge.argtypes = [pp, C.c_int64, dp, dp, dp, dp, dp, C.c_int, C.c_double, C.c_double]
        L.orc_particle_merge.restype = C.c_int64
        L.orc_particle_remove.argtypes = [C.c_int64, dp, dp, dp, dp, dp, C.c_int, C.c_double, C.c_double, C.c_int,
                                          C.c_double, C.c_void_p]
        L.orc_particle_remove.restype = C.c_int64
        L.orc_particle_split.argtypes = [pp, C.c_int64, C.c_int64, dp, dp, dp, dp, dp, C.c_int, C.c_double, C.c_double,
                                         C.c_int, C.c_double, C.c_void_p]
        L.orc_particle_split.restype = C.c_int64
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def default_params(**over) -> OrcParams:
    """Defaults of src/pic1dp_input.F90; keyword overrides set scalar fields or (for per-species / per-mode
    fields) a sequence."""
    p = OrcParams()
    lib().orc_params_default(C.byref(p))
    for k, v in over.items():
        cur = getattr(p, k)
        if hasattr(cur, "__len__"):
            for i, vi in enumerate(v):
                cur[i] = vi
        else:
            setattr(p, k, v)
    return p


class Oracle:
    """Thin object wrapper: holds parameters + field operators, exposes the restated subroutines."""

    def __init__(self, params: OrcParams):
        self.p = params
        self.L = lib()
        nx, M = params.nx, params.nmode
        self.F_re = np.zeros(nx * M)
        self.F_im = np.zeros(nx * M)
        self.grad_inv = np.zeros(M)
        self.L.orc_field_init(C.byref(self.p), _dp(self.F_re), _dp(self.F_im), _dp(self.grad_inv))

    # field_solve_electric
    def field_solve(self, rho: np.ndarray):
        nx, M = self.p.nx, self.p.nmode
        E, mre, mim = np.zeros(nx), np.zeros(M), np.zeros(M)
        self.L.orc_field_solve(C.byref(self.p), _dp(self.F_re), _dp(self.F_im), _dp(self.grad_inv),
                               _dp(np.ascontiguousarray(rho)), _dp(E), _dp(mre), _dp(mim))
        return E, mre, mim

    def shape(self, x: np.ndarray, right_frac: bool = False):
        n = x.size
        ix = np.zeros(n, dtype=np.int32)
        sl, sr = np.zeros(n), np.zeros(n)
        self.L.orc_shape(C.byref(self.p), n, _dp(x), ix.ctypes.data_as(C.POINTER(C.c_int32)), _dp(sl), _dp(sr),
                         int(right_frac))
        return ix, sl, sr

    def deposit_species(self, x: np.ndarray, w: np.ndarray):
        c1 = np.zeros(self.p.nx)
        noob = self.L.orc_deposit_species(C.byref(self.p), x.size, _dp(x), _dp(w), _dp(c1))
        return c1, noob

    # interaction_collect_charge over emulated ranks; xs/ws: list[species][rank] of arrays (x wrapped in place)
    def collect_charge(self, xs, ws):
        S, R = len(xs), len(xs[0])
        dpt = C.POINTER(C.c_double)
        npa = (C.c_int64 * (S * R))(*[xs[s][r].size for s in range(S) for r in range(R)])
        xa = (dpt * (S * R))(*[_dp(xs[s][r]) for s in range(S) for r in range(R)])
        wa = (dpt * (S * R))(*[_dp(ws[s][r]) for s in range(S) for r in range(R)])
        rho = np.zeros(self.p.nx)
        noob = self.L.orc_collect_charge(C.byref(self.p), R, npa, xa, wa, _dp(rho))
        return rho, noob

    # interaction_push_particle for one species on one rank (arrays updated in place)
    def push_species(self, isp, irk, x, v, p, w, xb, vb, wb, E):
        self.L.orc_push_species(C.byref(self.p), isp, irk, x.size, _dp(x), _dp(v), _dp(p), _dp(w), _dp(xb), _dp(vb),
                                _dp(wb), _dp(np.ascontiguousarray(E)))

    def dlnf0(self, isp, v):
        return self.L.orc_dlnf0(C.byref(self.p), isp, float(v))

    def field_energy(self, E):
        return self.L.orc_field_energy(C.byref(self.p), _dp(np.ascontiguousarray(E)))

    def _ptrs(self, states, key):
        S, R = len(states), len(states[0])
        dpt = C.POINTER(C.c_double)
        return (dpt * (S * R))(*[_dp(states[s][r][key]) for s in range(S) for r in range(R)])

    def output_field(self, states, E):
        """states: list[species][rank] of marker dicts.  Returns [energy, (sum v2, sum v2 p, sum v2 w | pert) ...]."""
        S, R = len(states), len(states[0])
        npa = (C.c_int64 * (S * R))(*[states[s][r]["x"].size for s in range(S) for r in range(R)])
        out = np.zeros(1 + 3 * S)
        self.L.orc_output_field(C.byref(self.p), R, npa, self._ptrs(states, "v"), self._ptrs(states, "p"),
                                self._ptrs(states, "w"), _dp(np.ascontiguousarray(E)), _dp(out))
        return out

    def output_ptcldist(self, states, isp, nx_opd=64, nv_opd=64, v_max=8.0):
        S, R = len(states), len(states[0])
        npa = (C.c_int64 * (S * R))(*[states[s][r]["x"].size for s in range(S) for r in range(R)])
        nc = nx_opd * nv_opd
        outs = [np.zeros(nc), np.zeros(nc), np.zeros(nc), np.zeros(nv_opd), np.zeros(nv_opd), np.zeros(nv_opd)]
        self.L.orc_output_ptcldist(C.byref(self.p), isp, R, npa, self._ptrs(states, "x"), self._ptrs(states, "v"),
                                   self._ptrs(states, "p"), self._ptrs(states, "w"), nx_opd, nv_opd, float(v_max),
                                   *[_dp(o) for o in outs])
        return dict(zip(("markr_xv", "total_xv", "pertb_xv", "markr_v", "total_v", "pertb_v"), outs))

    def particle_load(self, isp, al_int, mype, warmup, nlocal, ninit_total):
        x, v, pp, w = (np.zeros(nlocal) for _ in range(4))
        self.L.orc_particle_load(C.byref(self.p), isp, al_int, mype, warmup, nlocal, ninit_total, _dp(x), _dp(v),
                                 _dp(pp), _dp(w))
        return x, v, pp, w

    # ---- marker optimisation (src/pic1dp_particle.F90:356-746); arrays are updated in place, new np returned ----
    def dist_pertb_abs_v(self, vs, ws, nv=128, v_max=8.0):
        """vs, ws: per-rank arrays of one species.  Returns particle_dist_pertb_abs_v(ispecies, :)."""
        R = len(vs)
        dpt = C.POINTER(C.c_double)
        npa = (C.c_int64 * R)(*[a.size for a in vs])
        dist = np.zeros(nv)
        self.L.orc_dist_pertb_abs_v(R, npa, (dpt * R)(*[_dp(a) for a in vs]), (dpt * R)(*[_dp(a) for a in ws]), nv,
                                    float(v_max), _dp(dist))
        return dist

    def particle_merge(self, st, n, dist, thsh, v_max=8.0):
        return self.L.orc_particle_merge(C.byref(self.p), n, _dp(st["x"]), _dp(st["v"]), _dp(st["p"]), _dp(st["w"]),
                                         _dp(dist), dist.size, float(v_max), float(thsh))

    def particle_remove(self, st, n, dist, thsh, typeremove, remove_frac, rng, v_max=8.0):
        return self.L.orc_particle_remove(n, _dp(st["x"]), _dp(st["v"]), _dp(st["p"]), _dp(st["w"]), _dp(dist),
                                          dist.size, float(v_max), float(thsh), typeremove, float(remove_frac), rng.g)

    def particle_split(self, st, n, dist, thsh, ngroup, dv_sig_frac, rng, v_max=8.0):
        """st arrays are sized to the capacity (particle_ip_high - particle_ip_low)."""
        return self.L.orc_particle_split(C.byref(self.p), n, st["x"].size, _dp(st["x"]), _dp(st["v"]), _dp(st["p"]),
                                         _dp(st["w"]), _dp(dist), dist.size, float(v_max), float(thsh), ngroup,
                                         float(dv_sig_frac), rng.g)

    def run(self, states, nsteps, E0, nthreads=0):
        """states: list[species][rank] of dict(x,v,p,w) (updated in place).  Returns dict with rho,E,modes,
        energy trace and wall seconds; the loop is that of src/pic1dp.F90:78-93."""
        S, R = len(states), len(states[0])
        arr = (OrcRankState * (S * R))()
        keep = []
        for s in range(S):
            for r in range(R):
                st = states[s][r]
                n = st["x"].size
                for k in ("xb", "vb", "wb"):
                    st.setdefault(k, np.zeros(n))
                a = arr[s * R + r]
                a.np = n
                for k in ("x", "v", "p", "w", "xb", "vb", "wb"):
                    setattr(a, k, _dp(st[k]))
                keep.append(st)
        nx, M = self.p.nx, self.p.nmode
        rho, E, mre, mim = np.zeros(nx), np.array(E0, dtype=np.float64).copy(), np.zeros(M), np.zeros(M)
        en = np.zeros(max(nsteps, 1))
        secs = self.L.orc_run(C.byref(self.p), R, arr, nsteps, _dp(rho), _dp(E), _dp(mre), _dp(mim), _dp(en), nthreads)
        return dict(rho=rho, E=E, mode_re=mre, mode_im=mim, energy=en[:nsteps], seconds=secs)


def petsc_decide(n: int, npe: int, rank: int):
    lo, hi = C.c_int64(), C.c_int64()
    lib().orc_petsc_decide(n, npe, rank, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


class MultiRand:
    def __init__(self):
        self.L = lib()
        self.g = C.c_void_p(self.L.orc_multirand_new())

    def __del__(self):
        try:
            self.L.orc_multirand_free(self.g)
        except Exception:
            pass

    def seed_default(self, al_int):
        self.L.orc_multirand_seed_default(self.g, al_int)

    def init_const(self, al_int, mype, warmup):
        self.L.orc_multirand_init_const(self.g, al_int, mype, warmup)

    def int64(self):
        return self.L.orc_multirand_int64(self.g)

    def seeds4(self):
        out = (C.c_uint64 * 4)()
        self.L.orc_multirand_get_seeds4(self.g, out)
        return [int(v) for v in out]

    def skip(self, n):
        self.L.orc_multirand_skip(self.g, int(n))

    def real64(self):
        return self.L.orc_multirand_real64(self.g)

    def real_array(self, n):
        a = np.zeros(n)
        self.L.orc_multirand_real_array(self.g, _dp(a), n)
        return a

    def gaussian_array(self, n):
        a = np.zeros(n)
        self.L.orc_multirand_gaussian_array(self.g, _dp(a), n)
        return a
