"""On-device diagnostics against the oracle's restatement of output_field / output_ptcldist,
and the pic1dp.out writer against the py3 restatement of the reference's reader."""
import numpy as np
import pytest

import pic1dp_b200 as P
from helpers import OracleRun, copy_state, make_params, rel_err, synth_markers
from pic1dp_b200.output import OutputWriter
from tools_py3.output_data import OutputData

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["deltaf", "linear", "fullf_maxwell", "fullf_bump", "two_species"])
def test_output_field_and_ptcldist_match_oracle(case):
    kw, nsp = dict(nx=192, capacity=120000), 1
    if case == "linear":
        kw.update(linear=1)
    elif case == "fullf_maxwell":
        kw.update(deltaf=0, iptcldist=0, density=[1.0], v0=[0.5], temperature=[1.2])
    elif case == "fullf_bump":
        kw.update(deltaf=0)
    elif case == "two_species":
        nsp = 2
        kw.update(nspecies=2, charge=[-1.0, 1.0], mass=[1.0, 4.0], temperature=[1.0, 0.5], temperature2=[1.0, 0.5],
                  density=[0.9, 1.0], v0=[5.0, 0.0])
    op, gp = make_params(**kw)
    sts = [synth_markers(op, 120000 - 11 * s, seed=60 + s, isp=s) for s in range(nsp)]
    for st in sts:
        st["v"][:50] = np.linspace(-9.0, 9.0, 50)  # some |v| >= v_max markers must be skipped by the histogram
        st["x"][50] = op.lx                          # x == lx exactly
    ref = OracleRun(op, [[copy_state(s)] for s in sts])
    ref.init_field()
    ref.step()
    with P.Pic1dGpu(gp) as g:
        for s, st in enumerate(sts):
            g.set_markers(s, st["x"], st["v"], st["p"], st["w"])
        g.collect_charge()
        g.solve_field()
        g.step(1)
        sc = g.output_field()
        sc_ref = ref.o.output_field(ref.st, ref.E)
        # sums of positive terms: relative tolerance on each; the perturbed sums cancel -> relative to sum |terms|
        assert abs(sc[0] / sc_ref[0] - 1.0) < 1e-12
        for s in range(nsp):
            v, p, w = (ref.st[s][0][k] for k in ("v", "p", "w"))
            scales = [np.sum(v * v), np.sum(v * v * np.abs(p)), np.sum(v * v * np.abs(p)) if op.deltaf == 0 else np.sum(v * v * np.abs(w))]
            for q in range(3):
                assert abs(sc[1 + 3 * s + q] - sc_ref[1 + 3 * s + q]) < 1e-12 * scales[q], (case, s, q)
        for s in range(nsp):
            d = g.output_ptcldist(s, 64, 64, 8.0)
            dr = ref.o.output_ptcldist(ref.st, s, 64, 64, 8.0)
            for k in d:
                scale = np.max(np.abs(dr["total_xv" if k.endswith("xv") else "total_v"])) if "pertb" in k and op.deltaf == 0 \
                    else np.max(np.abs(dr[k]))
                assert rel_err(d[k], dr[k], scale) < 1e-11, (case, s, k)
            # the v-only histograms are the x-integrals of the x-v ones
            assert rel_err(d["markr_xv"].reshape(64, 64).sum(axis=1) * op.lx / 64, d["markr_v"]) < 1e-12


def test_ptcldist_odd_grid_sizes_and_repeat_calls():
    op, gp = make_params(nx=192, capacity=50000)
    st = synth_markers(op, 50000, seed=70)
    ref = OracleRun(op, [[copy_state(st)]])
    with P.Pic1dGpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        for nxo, nvo, vm in ((64, 64, 8.0), (16, 8, 4.0), (100, 33, 8.0), (64, 64, 8.0)):
            d = g.output_ptcldist(0, nxo, nvo, vm)
            dr = ref.o.output_ptcldist(ref.st, 0, nxo, nvo, vm)
            for k in d:
                assert rel_err(d[k], dr[k]) < 1e-11, (nxo, nvo, k)


def test_output_file_is_readable_by_the_reference_reader_layout(tmp_path):
    """100 steps of a quiet-start run written as pic1dp.out every 10 steps; the py3 OutputData restatement reads it
    back: header, scalars, fields, distributions; growth rate fit works on the file."""
    from test_gpu_physics import quiet_start
    op, gp = make_params(nx=192, capacity=400000)
    st = quiet_start(op, 500, 800)
    path = tmp_path / "pic1dp.out"
    with P.Pic1dGpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        g.collect_charge()
        g.solve_field()
        with OutputWriter(path, g) as out:
            out.output_all(0.0)
            for k in range(1, 61):
                g.step(10)
                last = out.output_all(0.5 * k)
        fld = g.get_field()
        h2d, d2h = g.counters().h2d_bytes, g.counters().d2h_bytes
    od = OutputData(str(path))
    assert (od.nspecies, od.nmode, od.nx, od.nx_pd, od.nv_pd) == (1, 1, 192, 64, 64)
    assert od.lx == op.lx and od.v_max == 8.0 and od.ntime == 61
    sc = od.get_scalar_t()
    assert np.allclose(sc[0], 0.5 * np.arange(61)) and sc[1, -1] == last[0]
    assert np.array_equal(od.get_field_x(60)[0, :192], fld["electric"])
    assert np.array_equal(od.get_mode_t()[1, -1:], fld["mode_im"])
    f = od.get_ptcldist_xv(60, 0, 1)
    assert f.shape == (64, 64) and abs(np.sum(f) * (op.lx / 64) * (16.0 / 63) - op.lx) < 0.02 * op.lx  # int f dx dv = n lx
    gamma = od.growthrate_energy_fit(15.0, 30.0) / 2.0
    assert abs(gamma / 0.0838311 - 1.0) < 0.05
    # no marker array ever crossed PCIe after the initial load
    assert d2h < 61 * (3 * 64 * 64 + 4 * 192 + 64) * 8 * 2


@pytest.mark.parametrize("case", ["deltaf", "fullf_maxwell", "two_species"])
def test_output_all_single_pass_matches_oracle(case):
    """pic1dp_gpu_output_all: the output_field sums and the x-v histograms of every species from ONE pass over the
    markers (fused kernel), against the oracle and against the two separate entry points."""
    kw, nsp = dict(nx=192, capacity=150000), 1
    if case == "fullf_maxwell":
        kw.update(deltaf=0, iptcldist=0, density=[1.0], v0=[0.5], temperature=[1.2])
    elif case == "two_species":
        nsp = 2
        kw.update(nspecies=2, charge=[-1.0, 1.0], mass=[1.0, 4.0], temperature=[1.0, 0.5], temperature2=[1.0, 0.5],
                  density=[0.9, 1.0], v0=[5.0, 0.0])
    op, gp = make_params(**kw)
    sts = [synth_markers(op, 150000 - 13 * s, seed=65 + s, isp=s) for s in range(nsp)]
    for st in sts:
        st["v"][:50] = np.linspace(-9.0, 9.0, 50)
        st["x"][50] = op.lx
        st["v"][51] = np.nextafter(8.0, 0.0)    # (v + v_max) / (2 v_max) rounds to 1: iv + 1 == nv_opd (spare row)
        st["v"][52] = -8.0                       # |v| >= v_max is skipped
    ref = OracleRun(op, [[copy_state(s)] for s in sts])
    ref.init_field()
    with P.Pic1dGpu(gp) as g:
        for s, st in enumerate(sts):
            g.set_markers(s, st["x"], st["v"], st["p"], st["w"])
        g.collect_charge()
        g.solve_field()
        sc, dists = g.output_all(64, 64, 8.0)
        sc_sep = g.output_field()
        sc2, dists2 = g.output_all(64, 64, 8.0)     # repeatable (private copies are cleared by the reduction)
        sc_ref = ref.o.output_field(ref.st, ref.E)
        assert np.allclose(sc, sc_ref, rtol=1e-11, atol=0) or np.max(np.abs(sc - sc_ref)) < 1e-12 * np.max(np.abs(sc_ref))
        assert np.max(np.abs(sc - sc_sep)) <= 1e-12 * np.max(np.abs(sc_sep))
        assert np.array_equal(sc[:1], sc2[:1])
        for s in range(nsp):
            dr = ref.o.output_ptcldist(ref.st, s, 64, 64, 8.0)
            dsep = g.output_ptcldist(s, 64, 64, 8.0)
            for k in dr:
                scale = np.max(np.abs(dr["total_xv" if k.endswith("xv") else "total_v"])) if "pertb" in k and op.deltaf == 0 \
                    else np.max(np.abs(dr[k]))
                # the oracle (like the reference) adds the iv + 1 row of the v ~ v_max marker out of bounds; its weight
                # 1 - sv is 0 there, so the arrays still agree
                assert rel_err(dists[s][k], dr[k], scale) < 1e-11, (case, s, k)
                assert rel_err(dists[s][k], dsep[k], scale) < 1e-11, (case, s, k)
                assert rel_err(dists2[s][k], dists[s][k], scale) < 1e-11, (case, s, k)


def test_output_all_odd_histogram_grids():
    op, gp = make_params(nx=192, capacity=60000)
    st = synth_markers(op, 60000, seed=71)
    ref = OracleRun(op, [[copy_state(st)]])
    with P.Pic1dGpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        for nxo, nvo, vm in ((64, 64, 8.0), (16, 8, 4.0), (100, 33, 8.0), (1, 2, 8.0), (300, 200, 8.0)):
            _, dists = g.output_all(nxo, nvo, vm)       # 300 x 200 does not fit in shared memory: RED fallback
            dr = ref.o.output_ptcldist(ref.st, 0, nxo, nvo, vm)
            for k in dr:
                assert rel_err(dists[0][k], dr[k]) < 1e-11, (nxo, nvo, k)


def _hist_all(gp, st, nxo=64, nvo=64, vm=8.0):
    with P.Pic1dGpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        a = g.output_all(nxo, nvo, vm)
        b = g.output_all(nxo, nvo, vm)
        c = g.output_ptcldist(0, nxo, nvo, vm)
    return a, b, c


def test_limb_histograms_are_bitwise_reproducible_and_match_the_cas_kernel(monkeypatch):
    """The default histogram kernel (k_diag_limb) accumulates exact fixed-point integers with native 32-bit shared-memory
    adds: the result does not depend on the arrival order, so repeated calls and a second handle agree bit for bit;
    against the round-2 CAS.128 kernel (PIC1DP_DIAG_CAS=1, fp64 sums in arrival order) it agrees to 1e-12 of the maximum."""
    op, gp = make_params(nx=192, capacity=300000)
    st = synth_markers(op, 300000, seed=72)
    st["w"][::7] *= -1.0
    (sc1, d1), (sc2, d2), sep1 = _hist_all(gp, st)
    (sc3, d3), _, _ = _hist_all(gp, st)
    for k in d1[0]:
        assert np.array_equal(d1[0][k], d2[0][k]), k      # same handle, second call
        assert np.array_equal(d1[0][k], d3[0][k]), k      # another handle
        assert np.array_equal(d1[0][k], sep1[k]), k       # output_ptcldist runs the same kernel without the sums
    monkeypatch.setenv("PIC1DP_DIAG_CAS", "1")
    (sc4, d4), _, _ = _hist_all(gp, st)
    monkeypatch.delenv("PIC1DP_DIAG_CAS")
    assert np.array_equal(sc1, sc4)                       # the output_field sums do not go through the histogram
    for k in d1[0]:
        assert rel_err(d1[0][k], d4[0][k]) < 1e-12, k


@pytest.mark.parametrize("case", ["outlier", "zero_w", "tiny", "huge", "one_cell"])
def test_limb_histograms_dynamic_range(case):
    """The fixed-point scale comes from the largest |p| and |w| of the species: one huge outlier costs the other markers
    resolution (2^-42 of the outlier), all-zero and very small / very large weights must not break the scaling, and the
    worst case of the counter bounds (every marker in one cell, same sign) stays exact."""
    op, gp = make_params(nx=192, capacity=200000)
    st = synth_markers(op, 200000, seed=73)
    if case == "outlier":
        st["w"][1234] = 1.0e3 * np.max(np.abs(st["w"]))
        st["p"][4321] = 1.0e3 * np.max(np.abs(st["p"]))
    elif case == "zero_w":
        st["w"][:] = 0.0
    elif case == "tiny":
        st["w"] *= 1e-200
        st["p"] *= 1e-100
    elif case == "huge":
        st["w"] *= 1e250
        st["p"] *= 1e200
    elif case == "one_cell":
        st["x"][:] = 0.4321 * op.lx / 64 + 10 * op.lx / 64
        st["v"][:] = 0.7
        st["w"][:] = np.abs(st["w"]).max()
        st["p"][:] = np.abs(st["p"]).max()
    ref = OracleRun(op, [[copy_state(st)]])
    (sc, d), _, _ = _hist_all(gp, st)
    dr = ref.o.output_ptcldist(ref.st, 0, 64, 64, 8.0)
    # one_cell: the oracle's own sequential fp64 sum of 2e5 equal terms is only good to ~1e-11; the integer sum is exact
    tol = 1e-10 if case == "one_cell" else 1e-11
    for k in dr:
        scale = np.max(np.abs(dr[k]))
        if scale == 0.0:
            assert not np.any(d[0][k]), (case, k)
        else:
            assert rel_err(d[0][k], dr[k], scale) < tol, (case, k)
    if case == "one_cell":
        cell = (op.lx / 64) * (16.0 / 63)
        assert abs(np.sum(d[0]["markr_xv"]) * cell / st["x"].size - 1.0) < 1e-12
        assert abs(np.sum(d[0]["pertb_xv"]) * cell / (st["x"].size * st["w"][0]) - 1.0) < 1e-12


def test_limb_histograms_counter_flush_at_scale():
    """2.2e7 markers: every CTA runs more than PIC1DP_LIMB_FLUSH = 128 tile steps, so its 32-bit counters are folded into
    the 64-bit table in mid-kernel; conservation (sum of the g histogram = markers inside |v| < v_max, sum of f / delta f
    = sum of p / w over them) holds to rounding and the oracle agrees on a strided sample of the same markers."""
    n = 22_000_000
    op, gp = make_params(nx=1024, capacity=n)
    st = synth_markers(op, n, seed=74)
    with P.Pic1dGpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        sc, d = g.output_all(64, 64, 8.0)
        sc2, d2 = g.output_all(64, 64, 8.0)
    inside = np.abs(st["v"]) < 8.0
    cell = (op.lx / 64) * (16.0 / 63)      # the reference scales the x-v histograms by 1 / (delx delv)
    assert abs(np.sum(d[0]["markr_xv"]) * cell / np.count_nonzero(inside) - 1.0) < 1e-12
    assert abs(np.sum(d[0]["total_xv"]) * cell / np.sum(st["p"][inside]) - 1.0) < 1e-12
    assert abs(np.sum(d[0]["pertb_xv"]) * cell - np.sum(st["w"][inside])) < 1e-12 * np.sum(np.abs(st["w"][inside]))
    for k in d[0]:
        assert np.array_equal(d[0][k], d2[0][k]), k
    ref = OracleRun(op, [[{k: st[k].copy() for k in ("x", "v", "p", "w")}]])
    dr = ref.o.output_ptcldist(ref.st, 0, 64, 64, 8.0)
    for k in dr:
        assert rel_err(d[0][k], dr[k]) < 1e-11, k


@pytest.mark.parametrize("fuse", [1, 0])
def test_limb_histogram_scale_follows_the_markers(fuse):
    """The fixed-point scale of the delta-f histogram comes from a device-resident bound on max |w|: the running maximum
    that the (default) fixed-point deposit keeps in fused runs, a fresh pass otherwise.  Whatever the call sequence, the
    histograms must be those of the markers the device holds at that moment (oracle binning of get_markers())."""
    op, gp = make_params(nx=256, capacity=150000, fuse=fuse)
    st = synth_markers(op, 150000, seed=75)
    ref = OracleRun(op, [[copy_state(st)]])

    def check(g, what):
        mk = g.get_markers(0)
        d = g.output_ptcldist(0, 64, 64, 8.0)
        dr = ref.o.output_ptcldist([[{k: mk[k] for k in ("x", "v", "p", "w")}]], 0, 64, 64, 8.0)
        for k in dr:
            assert rel_err(d[k], dr[k]) < 1e-11, (what, k)

    with P.Pic1dGpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        check(g, "fresh markers, nothing deposited yet")
        g.collect_charge()
        g.solve_field()
        g.step(3)
        check(g, "after three steps")
        g.push(1)                                  # w has moved since the last output
        if fuse:                                   # (unfused, x is not wrapped before collect_charge: nothing to bin yet)
            check(g, "between push and collect_charge")
        g.collect_charge()
        g.solve_field()
        big = {k: st[k].copy() for k in st}
        big["w"] *= 1.0e4                          # replaced markers with a much larger |w| than any bound kept so far
        g.set_markers(0, big["x"], big["v"], big["p"], big["w"])
        check(g, "after set_markers with larger weights")
        small = {k: st[k].copy() for k in st}
        small["w"] *= 1.0e-6
        g.set_markers(0, small["x"], small["v"], small["p"], small["w"])
        check(g, "after set_markers with smaller weights")
