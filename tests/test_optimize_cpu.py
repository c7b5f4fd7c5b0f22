"""Marker optimisation (/root/reference/src/pic1dp_particle.F90:356-813), host halves, no GPU needed:
the product's particle_merge / particle_remove / particle_split (libpic1dp_b200.so, pic1dp_host_*) against the
oracle's statement-by-statement restatement on identical marker arrays and identical multirand streams -- bit for bit,
including the marker ORDER the swap-with-last bookkeeping leaves behind."""
import ctypes as C

import numpy as np
import pytest

from helpers import make_params, synth_markers
from oracle import oracle as O
from pic1dp_b200 import _capi

NV, VMAX = 128, 8.0


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _markers(n, seed, nx=192, cap=None, grow_w=True):
    op, _ = make_params(nx=nx)
    st = synth_markers(op, n, seed=seed, spread=0.3)
    if grow_w:  # a developed perturbation: resonant markers carry most of |w|
        st["w"] = st["w"] * (1.0 + 50.0 * np.exp(-(st["v"] - 3.2) ** 2)) * np.sign(np.sin(7.0 * st["x"]) + 0.3)
    if cap:
        for k in st:
            st[k] = np.concatenate([st[k], np.full(cap - n, np.nan)])
    return op, st


def _dist(op, st, n):
    return O.Oracle(op).dist_pertb_abs_v([st["v"][:n].copy()], [st["w"][:n].copy()], NV, VMAX)


def _same(a, b, n):
    for k in ("x", "v", "p", "w"):
        assert np.array_equal(a[k][:n], b[k][:n]), k


def test_dist_pertb_abs_v_definition():
    """Oracle's particle_compute_dist_pertb_abs_v against a direct numpy statement of :378-390."""
    op, st = _markers(20000, 3)
    d = _dist(op, st, 20000)
    v, w = st["v"], st["w"]
    ok = np.abs(v) < VMAX
    sv = (v[ok] + VMAX) / (VMAX * 2.0) * (NV - 1)
    iv = np.floor(sv).astype(int)
    s = 1.0 - (sv - iv)
    ref = np.zeros(NV + 1)
    np.add.at(ref, iv, s * np.abs(w[ok]))
    np.add.at(ref, iv + 1, (1.0 - s) * np.abs(w[ok]))
    assert np.allclose(d, ref[:NV], rtol=1e-12, atol=0) and d.max() > 0
    two = O.Oracle(op).dist_pertb_abs_v([v[:9000].copy(), v[9000:].copy()], [w[:9000].copy(), w[9000:].copy()], NV, VMAX)
    assert np.allclose(two, d, rtol=1e-13)  # MPI_Allreduce over two emulated ranks


@pytest.mark.parametrize("n,nx,thsh,seed", [(50000, 192, 0.1, 1), (50000, 16, 0.5, 2), (3000, 8, 2.0, 3), (1, 192, 0.1, 4),
                                            (0, 192, 0.1, 5), (2, 4, 2.0, 6)])
def test_merge_matches_oracle_bit_for_bit(n, nx, thsh, seed):
    op, a = _markers(n, seed, nx=nx)
    b = {k: v.copy() for k, v in a.items()}
    dist = _dist(op, a, n) if n else np.ones(NV)
    n_ref = O.Oracle(op).particle_merge(a, n, dist, thsh, VMAX)
    n_got = _capi.load().pic1dp_host_particle_merge(n, _dp(b["x"]), _dp(b["v"]), _dp(b["p"]), _dp(b["w"]), _dp(dist), NV,
                                                    VMAX, thsh, nx, op.lx)
    assert n_got == n_ref
    _same(a, b, n_ref)
    if n >= 3000:
        assert n_ref < n, "the case must actually merge something"
        # merging conserves sum p and sum w of the examined species (p1 + p, w1 + w :492-493) up to rounding
        assert abs(b["p"][:n_got].sum() / a["p"][:n_ref].sum() - 1) < 1e-12


@pytest.mark.parametrize("n,typeremove,thsh,frac,seed", [(40000, 2, 0.0, 0.9, 1), (40000, 1, 0.2, 0.9, 2),
                                                          (40000, 1, 5.0, 0.5, 3), (1, 2, 0.1, 0.9, 4), (0, 2, 0.1, 0.9, 5)])
def test_remove_matches_oracle_bit_for_bit(n, typeremove, thsh, frac, seed):
    op, a = _markers(n, seed)
    b = {k: v.copy() for k, v in a.items()}
    dist = _dist(op, a, n) if n else np.ones(NV)
    r1, r2 = O.MultiRand(), O.MultiRand()
    r1.init_const(3, 0, 5)
    r2.init_const(3, 0, 5)
    n_ref = O.Oracle(op).particle_remove(a, n, dist, thsh, typeremove, frac, r1, VMAX)
    cb = _capi.REAL64_FN(lambda _ctx: r2.real64())  # the same stream, drawn through the product's call-back
    n_got = _capi.load().pic1dp_host_particle_remove(n, _dp(b["x"]), _dp(b["v"]), _dp(b["p"]), _dp(b["w"]), _dp(dist), NV,
                                                     VMAX, thsh, typeremove, frac, cb, None)
    assert n_got == n_ref
    _same(a, b, n_ref)
    assert r1.int64() == r2.int64(), "both consumed the same number of draws"
    if n >= 40000:
        assert 0 < n_ref < n


@pytest.mark.parametrize("n,cap,thsh,ngroup,deltaf,seed", [(20000, 120000, 0.5, 5, 1, 1), (20000, 20100, 0.1, 5, 1, 2),
                                                           (5000, 5008, 0.0, 5, 1, 3), (5000, 60000, 0.3, 1, 1, 4),
                                                           (5000, 60000, 0.3, 3, 0, 5)])
def test_split_matches_oracle_bit_for_bit(n, cap, thsh, ngroup, deltaf, seed):
    op, a = _markers(n, seed, cap=cap)
    op.deltaf = deltaf
    b = {k: v.copy() for k, v in a.items()}
    dist = _dist(op, a, n)
    r1, r2 = O.MultiRand(), O.MultiRand()
    r1.init_const(3, 1, 5)
    r2.init_const(3, 1, 5)
    n_ref = O.Oracle(op).particle_split(a, n, dist, thsh, ngroup, 0.1, r1, VMAX)

    def fill(_ctx, arr, k):
        g = r2.gaussian_array(k)
        for i in range(k):
            arr[i] = g[i]
    cb = _capi.GAUSSIAN_ARRAY_FN(fill)
    n_got = _capi.load().pic1dp_host_particle_split(n, cap, _dp(b["x"]), _dp(b["v"]), _dp(b["p"]), _dp(b["w"]), _dp(dist),
                                                    NV, VMAX, thsh, ngroup, 0.1, deltaf, cb, None)
    assert n_got == n_ref and n <= n_ref <= cap
    for k in ("x", "v", "p") + (("w",) if deltaf else ()):
        assert np.array_equal(a[k][:n_ref], b[k][:n_ref]), k
    assert r1.int64() == r2.int64()
    if cap - n >= 2 * ngroup - 1 and thsh < 1.0:
        assert n_ref > n and (n_ref - n) % (2 * ngroup - 1) == 0


def test_split_shares_the_parent_weight():
    """2*ngroup children each carry 1/(2*ngroup) of the parent's p and w and sit at v +- dv (:706-728)."""
    op, a = _markers(2000, 9, cap=40000)
    before = {k: v[:2000].copy() for k, v in a.items()}
    dist = _dist(op, a, 2000)
    rng = O.MultiRand()
    rng.init_const(3, 0, 5)
    n2 = O.Oracle(op).particle_split(a, 2000, dist, 0.5, 5, 0.1, rng, VMAX)
    assert abs(a["p"][:n2].sum() - before["p"].sum()) < 1e-12 * abs(before["p"]).sum()
    assert abs(a["w"][:n2].sum() - before["w"].sum()) < 1e-12 * abs(before["w"]).sum()
    assert abs((a["v"][:n2] * a["p"][:n2]).sum() - (before["v"] * before["p"]).sum()) < 1e-9  # +dv and -dv cancel


def test_randomized_merge_remove_split_sequences_match_oracle():
    """40 pseudo-random (size, grid, threshold, mode) draws, each applying merge -> remove -> split in sequence to the
    same arrays on both sides with one shared RNG stream position: marker counts and arrays stay bit-identical."""
    rs = np.random.default_rng(12345)
    L = _capi.load()
    for trial in range(40):
        n = int(rs.choice([0, 1, 2, 7, 300, 5000, 20000]))
        nx = int(rs.choice([2, 5, 64, 192]))
        nv = int(rs.choice([2, 16, 128]))
        cap = n + int(rs.choice([0, 3, 9, 4 * n + 50]))
        thm, ths = float(rs.choice([0.0, 0.05, 0.5, 2.0])), float(rs.choice([0.0, 0.2, 0.9, 1.5]))
        typeremove, frac = int(rs.choice([1, 2])), float(rs.choice([0.1, 0.5, 0.9]))
        ngroup = int(rs.choice([1, 2, 5]))
        op, a = _markers(n, 100 + trial, nx=nx, cap=max(cap, 1))
        b = {k: v.copy() for k, v in a.items()}
        orc = O.Oracle(op)
        r1, r2 = O.MultiRand(), O.MultiRand()
        r1.init_const(1 + trial % 3, trial, 2)
        r2.init_const(1 + trial % 3, trial, 2)
        na = nb = n
        cb_r = _capi.REAL64_FN(lambda _ctx: r2.real64())

        def fill(_ctx, arr, k):
            g = r2.gaussian_array(k)
            for i in range(k):
                arr[i] = g[i]
        cb_g = _capi.GAUSSIAN_ARRAY_FN(fill)
        for stage in ("merge", "remove", "split"):
            dist = orc.dist_pertb_abs_v([a["v"][:na].copy()], [a["w"][:na].copy()], nv, VMAX) if na else np.ones(nv)
            if not dist.max() > 0:
                dist = np.ones(nv)
            if stage == "merge":
                na = orc.particle_merge(a, na, dist, thm, VMAX)
                nb = L.pic1dp_host_particle_merge(nb, _dp(b["x"]), _dp(b["v"]), _dp(b["p"]), _dp(b["w"]), _dp(dist), nv, VMAX,
                                                  thm, nx, op.lx)
            elif stage == "remove":
                na = orc.particle_remove(a, na, dist, thm, typeremove, frac, r1, VMAX)
                nb = L.pic1dp_host_particle_remove(nb, _dp(b["x"]), _dp(b["v"]), _dp(b["p"]), _dp(b["w"]), _dp(dist), nv, VMAX,
                                                   thm, typeremove, frac, cb_r, None)
            else:
                na = orc.particle_split(a, na, dist, ths, ngroup, 0.1, r1, VMAX)
                nb = L.pic1dp_host_particle_split(nb, a["x"].size, _dp(b["x"]), _dp(b["v"]), _dp(b["p"]), _dp(b["w"]),
                                                  _dp(dist), nv, VMAX, ths, ngroup, 0.1, 1, cb_g, None)
            assert na == nb, (trial, stage)
            for k in ("x", "v", "p", "w"):
                assert np.array_equal(a[k][:na], b[k][:na], equal_nan=True), (trial, stage, k)
        assert r1.int64() == r2.int64(), trial


def _golden():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(__file__), "golden", "optimize_tiny.json")))


def _unhex(a):
    return np.array([float.fromhex(t) for t in a])


@pytest.mark.parametrize("side", ["oracle", "product"])
def test_committed_optimisation_fixture_is_reproduced(side):
    """tests/golden/optimize_tiny.json (oracle-generated, tests/golden/make_golden.py): merge -> remove -> split on 400
    loader markers; the oracle (regression) and the product's host halves (given the fixture's dist, no oracle
    arithmetic) reproduce every stage bit for bit and leave the RNG stream at the same position."""
    g = _golden()
    nx, nv, vmax, cap, m = g["nx"], g["nv"], g["v_max"], g["capacity"], g["n0"]
    op, _ = make_params(nx=nx)
    st = {k: np.concatenate([_unhex(g["init"][k]), np.zeros(cap - m)]) for k in ("x", "v", "p", "w")}
    rng = O.MultiRand()
    rng.init_const(g["rng"]["al_int"], g["rng"]["mype"], g["rng"]["warmup"])
    L = _capi.load()
    cb_r = _capi.REAL64_FN(lambda _ctx: rng.real64())

    def fill(_ctx, arr, k):
        gz = rng.gaussian_array(k)
        for i in range(k):
            arr[i] = gz[i]
    cb_g = _capi.GAUSSIAN_ARRAY_FN(fill)
    orc = O.Oracle(op)
    for rec in g["stages"]:
        dist = _unhex(rec["dist"])
        if side == "oracle":
            assert np.array_equal(orc.dist_pertb_abs_v([st["v"][:m].copy()], [st["w"][:m].copy()], nv, vmax), dist)
            if rec["stage"] == "merge":
                m = orc.particle_merge(st, m, dist, 0.5, vmax)
            elif rec["stage"] == "remove":
                m = orc.particle_remove(st, m, dist, 0.0, 2, 0.9, rng, vmax)
            else:
                m = orc.particle_split(st, m, dist, 0.5, 2, 0.1, rng, vmax)
        else:
            a = [_dp(st[k]) for k in ("x", "v", "p", "w")]
            if rec["stage"] == "merge":
                m = L.pic1dp_host_particle_merge(m, *a, _dp(dist), nv, vmax, 0.5, nx, op.lx)
            elif rec["stage"] == "remove":
                m = L.pic1dp_host_particle_remove(m, *a, _dp(dist), nv, vmax, 0.0, 2, 0.9, cb_r, None)
            else:
                m = L.pic1dp_host_particle_split(m, cap, *a, _dp(dist), nv, vmax, 0.5, 2, 0.1, 1, cb_g, None)
        assert m == rec["np"], rec["stage"]
        for k in ("x", "v", "p", "w"):
            assert np.array_equal(st[k][:m], _unhex(rec[k])), (rec["stage"], k)
    assert rng.int64() == g["next_int64"]
