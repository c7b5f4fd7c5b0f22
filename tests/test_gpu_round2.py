"""Round-2 features through the C ABI: fixed-point deterministic deposit (all grid sizes), the replayed step graph,
device-side random streams for particle_load, TOLERANCE arithmetic, and error surfacing."""
import numpy as np
import pytest

import pic1dp_b200 as P
from oracle import oracle as O
from helpers import OracleRun, copy_state, make_params, rel_err, synth_markers

pytestmark = pytest.mark.gpu

TOL_SUM = 1e-12


def _run(gp, st, nsteps):
    with P.Pic1dGpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        g.collect_charge()
        g.solve_field()
        g.step(nsteps)
        return g.get_field(), g.get_markers(0), g.counters()


@pytest.mark.parametrize("nx,req", [(1024, P.DEPOSIT_FIXED), (4096, P.DEPOSIT_FIXED), (8192, P.DEPOSIT_FIXED),
                                    (4096, P.DEPOSIT_WARP_PRIVATE), (8192, P.DEPOSIT_WARP_PRIVATE)])
def test_deterministic_deposit_on_large_grids(nx, req):
    """Bitwise-identical rho / E / w across three runs at nx = 4096 and 8192 (configs[2], configs[4]), and parity with
    the oracle.  A WARP_PRIVATE request that does not fit degrades to FIXED instead of failing."""
    n = 600_001
    op, gp = make_params(nx=nx, capacity=n, deposit_mode=req)
    st = synth_markers(op, n, seed=21, spread=0.3)
    ref = OracleRun(op, [[copy_state(st)]])
    ref.init_field()
    for _ in range(4):
        ref.step()
    res = [_run(gp, st, 4) for _ in range(3)]
    assert res[0][2].deposit_mode == P.DEPOSIT_FIXED
    for f, mk, _ in res[1:]:
        assert np.array_equal(f["chargeden"], res[0][0]["chargeden"])
        assert np.array_equal(f["electric"], res[0][0]["electric"])
        assert np.array_equal(mk["w"], res[0][1]["w"]) and np.array_equal(mk["x"], res[0][1]["x"])
    f, mk, _ = res[0]
    assert rel_err(f["chargeden"], ref.rho) < TOL_SUM and rel_err(f["electric"], ref.E) < TOL_SUM
    for k in ("x", "v", "w"):
        assert rel_err(mk[k], ref.st[0][0][k]) < 1e-12, k


def test_fixed_deposit_is_independent_of_the_launch_geometry():
    """Integer sums do not depend on which warp adds first and the per-CTA grids are exact, so standalone deposits
    of the same markers agree to the last bit with the fused kernel's deposit of the same x, w."""
    n = 300_000
    op, gp = make_params(nx=2048, capacity=n, deposit_mode=P.DEPOSIT_FIXED)
    st = synth_markers(op, n, seed=22)
    rhos = []
    for fuse in (1, 0):
        gp.fuse = fuse
        with P.Pic1dGpu(gp) as g:
            g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
            g.collect_charge()
            g.solve_field()
            for irk in (1, 2):
                g.push(irk)
                g.collect_charge()
                g.solve_field()
            rhos.append(g.get_field()["chargeden"])
    assert np.array_equal(rhos[0], rhos[1])


def test_fixed_deposit_cannot_overflow_when_all_markers_share_a_cell():
    """The scale is chosen so that a slot holds the sum even if every marker of a CTA lands in ONE cell: put 3/4 of
    4.5e6 markers into a single cell (30 000 per CTA) and compare with the oracle."""
    n = 148 * 2048 * 15 + 777
    op, gp = make_params(nx=512, capacity=n, deposit_mode=P.DEPOSIT_FIXED)
    st = synth_markers(op, n, seed=23)
    st["x"][: 3 * n // 4] = 0.37 * op.lx
    st["w"][: 3 * n // 4] = np.abs(st["w"][: 3 * n // 4]).max()      # all maximal and of one sign: no cancellation
    ref = OracleRun(op, [[copy_state(st)]])
    ref.init_field()
    ref.step()
    a = _run(gp, st, 1)
    b = _run(gp, st, 1)
    for k in ("chargeden", "electric"):
        assert np.array_equal(a[0][k], b[0][k]), k
    assert rel_err(a[0]["chargeden"], ref.rho) < TOL_SUM and rel_err(a[0]["electric"], ref.E) < TOL_SUM


def test_fixed_deposit_reports_overflow_instead_of_wrong_density():
    """A field that makes w grow by many orders of magnitude inside one substep exceeds the fixed-point range: the next
    synchronising call must fail (PIC1DP_ESTATE), not return a silently wrong rho."""
    n = 50_000
    op, gp = make_params(nx=256, capacity=n, deposit_mode=P.DEPOSIT_FIXED)
    st = synth_markers(op, n, seed=24)
    with P.Pic1dGpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        g.collect_charge()
        g.solve_field()
        g.set_field(electric=np.full(op.nx, 1e9))
        g.push(1)
        g.collect_charge()
        with pytest.raises(P.Pic1dpError) as e:
            g.get_field()
        assert e.value.code == 5 and "overflow" in str(e.value)


@pytest.mark.parametrize("dep", [P.DEPOSIT_AUTO, P.DEPOSIT_SMEM_ATOMIC, P.DEPOSIT_FIXED, P.DEPOSIT_GLOBAL_RED])
def test_step_graph_replay_equals_direct_launches(dep):
    """pic1dp_gpu_step replays one captured graph per timestep; the result must be that of the individual launches,
    also when step() is called repeatedly, after the markers changed size, and mixed with per-call substeps."""
    op, gp = make_params(nx=512, capacity=120_000, deposit_mode=dep)
    st = synth_markers(op, 120_000, seed=25)
    outs = []
    for no_graph in (0, 1):
        gp.no_step_graph = no_graph
        with P.Pic1dGpu(gp) as g:
            g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
            g.collect_charge()
            g.solve_field()
            g.step(2)
            g.step(1)
            for irk in (1, 2):             # individual calls between graph replays
                g.push(irk)
                g.collect_charge()
                g.solve_field()
            g.step(2)
            mk = g.get_markers(0)
            g.set_markers(0, mk["x"][:70_001], mk["v"][:70_001], mk["p"][:70_001], mk["w"][:70_001])   # np changed
            g.collect_charge()
            g.solve_field()
            g.step(2)
            c = g.counters()
            assert (c.graph_replays == 7) if no_graph == 0 else (c.graph_replays == 0)
            outs.append((g.get_field(), g.get_markers(0)))
    if dep in (P.DEPOSIT_FIXED,):   # deterministic deposit: bitwise
        assert np.array_equal(outs[0][0]["electric"], outs[1][0]["electric"])
        assert np.array_equal(outs[0][1]["w"], outs[1][1]["w"])
    else:
        assert rel_err(outs[0][0]["electric"], outs[1][0]["electric"]) < 1e-11
        assert rel_err(outs[0][1]["w"], outs[1][1]["w"]) < 1e-11
    assert rel_err(outs[0][1]["x"], outs[1][1]["x"]) < 1e-12


def test_step_graph_against_oracle_small_default_problem():
    op, gp = make_params(nx=192, capacity=200_000)
    st = synth_markers(op, 200_000, seed=26)
    ref = OracleRun(op, [[copy_state(st)]])
    ref.init_field()
    for _ in range(6):
        ref.step()
    f, mk, c = _run(gp, st, 6)
    assert c.graph_replays == 6
    assert rel_err(f["chargeden"], ref.rho) < TOL_SUM and rel_err(f["electric"], ref.E) < TOL_SUM
    for k in ("x", "v", "w"):
        assert rel_err(mk[k], ref.st[0][0][k]) < 1e-12, k


# ---- device-side random streams ----

DEFAULT_SEEDS = [1234567890987654321, 362436362436362436, 1066149217761810, 123456123456123456]


def test_device_kiss64_known_answer_and_offsets():
    """The device stream equals the reference's known-answer vector (src/multirand.F90:396-401) and the KAT-pinned
    oracle's sequential stream at offsets 0, 1e6 and 1e8."""
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "multirand_kat.json")))["kiss64"]
    op, gp = make_params(capacity=16)
    with P.Pic1dGpu(gp) as g:
        u = g.kiss64_uniforms(DEFAULT_SEEDS, 0, 10)
        assert [float(t) for t in u] == [float(np.int64(i)) / 18446744073709551615.0 + 0.5 for i in gold]
        for mype, offset in ((0, 0), (2, 1_000_000), (1, 100_000_000)):
            r = O.MultiRand()
            r.init_const(1, mype, 5)
            seeds = r.seeds4()
            r.skip(offset)
            want = r.real_array(200_003)       # several 512-number chunks per thread + a ragged tail
            got = g.kiss64_uniforms(seeds, offset, 200_003)
            assert np.array_equal(got, want), (mype, offset)


@pytest.mark.parametrize("dist", [0, 3])
def test_device_kiss64_particle_load_matches_oracle_loader(dist):
    """particle_load with multirand_al_int = 1 (KISS64), seed_type 1: the device generates both uniform streams; x, v
    bit-exact against the oracle's loader, p, w to a few ulp (exp / sin on the device)."""
    n = 300_007
    kw = dict(nx=256, capacity=n, iptcldist=dist)
    if dist == 0:
        kw.update(density=[1.0], v0=[0.2])
    op, gp = make_params(**kw)
    o = O.Oracle(op)
    for mype in (0, 3):
        x, v, p, w = o.particle_load(0, 1, mype, 5, n, n)      # al_int = 1
        r = O.MultiRand()
        r.init_const(1, mype, 5)
        with P.Pic1dGpu(gp) as g:
            g.load_markers_kiss64(0, n, r.seeds4(), 0, n, n, v_max=op.v_max)   # pv draws first, then px (:180, :222)
            out = g.get_markers(0)
            assert g.counters().h2d_bytes < 1_000_000     # no marker-sized upload
        assert np.array_equal(out["x"], x) and np.array_equal(out["v"], v)
        assert rel_err(out["p"], p) < 1e-14 and rel_err(out["w"], w) < 1e-13


def test_counter_based_load_is_decomposition_independent():
    """load_markers_counter: the uniforms depend on (seed, species, global index) only -- two ranks' halves equal one
    rank's whole, and the device stream equals the host's evaluation of the same generator."""
    from pic1dp_b200 import host as H
    n = 100_001
    op, gp = make_params(nx=256, capacity=n)
    with P.Pic1dGpu(gp) as g:
        g.load_markers_counter(0, n, 77, 0, n)
        whole = g.get_markers(0)
        g.load_markers_counter(0, 40_000, 77, 0, n)
        a = g.get_markers(0)
        g.load_markers_counter(0, n - 40_000, 77, 40_000, n)
        b = g.get_markers(0)
    for k in ("x", "v", "p", "w"):
        assert np.array_equal(np.concatenate([a[k], b[k]]), whole[k]), k
    u_v, u_x = H.host_counter_uniforms(77, 0, 0, n)
    assert np.array_equal(whole["x"], u_x * op.lx) and np.array_equal(whole["v"], (u_v - 0.5) * 2.0 * op.v_max)


# ---- PIC1DP_ARITH_TOLERANCE ----

@pytest.mark.parametrize("dist,consts", [(3, "unit"), (3, "nonpow2"), (3, "pow2"), (2, "unit"), (2, "nonpow2")])
@pytest.mark.parametrize("dep", [P.DEPOSIT_SMEM_ATOMIC, P.DEPOSIT_FIXED])
def test_tolerance_arithmetic_substeps_against_oracle(dist, consts, dep):
    """arith_mode = TOLERANCE (one exponential of the argument difference, fused multiply-adds on the w path): with a
    prescribed E both RK substeps give x, v BIT-EXACT (that path keeps the reference's operation order) and w within
    1e-14 of max|w| of the strict oracle -- the same bar the STRICT mode meets (its exp is not glibc's either)."""
    kw = dict(nx=256, iptcldist=dist, capacity=200_001, deposit_mode=dep, arith_mode=P.ARITH_TOLERANCE)
    if consts == "nonpow2":
        kw.update(temperature=[1.3], temperature2=[0.7], mass=[0.9])
    elif consts == "pow2":
        kw.update(temperature=[2.0], temperature2=[0.5], mass=[0.5])
    op, gp = make_params(**kw)
    n = 200_001
    st = synth_markers(op, n, seed=40 + dist)
    E = 1e-3 * np.sin(2 * np.pi * np.arange(op.nx) / op.nx + 0.3)
    ref = OracleRun(op, [[copy_state(st)]])
    ref.E = E.copy()
    with P.Pic1dGpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        g.set_field(electric=E)
        for irk in (1, 2):
            ref.push(irk)
            g.push(irk)
            out = g.get_markers(0)
            r = ref.st[0][0]
            xr = r["x"].copy()
            ref.o.shape(xr)   # the fused kernel also wraps
            assert np.array_equal(out["x"], xr) and np.array_equal(out["v"], r["v"]), irk
            assert rel_err(out["w"], r["w"]) < 1e-14, irk
            r["x"][:] = xr            # the reference wraps here too (collect_charge) before the next push
            r["w"][:] = out["w"]      # continue from identical state so irk = 2 isolates one substep


@pytest.mark.parametrize("dep", [P.DEPOSIT_AUTO, P.DEPOSIT_FIXED])
def test_tolerance_arithmetic_ten_steps_within_north_star_bar(dep):
    """Ten full steps of the default problem in TOLERANCE mode against the strict oracle: rho, E <= 1e-12 of max at
    every substep, x, v, w <= 1e-12 at the end, cell indices identical at every substep."""
    n = 400_000
    op, gp = make_params(nx=256, capacity=n, deposit_mode=dep, arith_mode=P.ARITH_TOLERANCE)
    st = synth_markers(op, n, seed=45)
    ref = OracleRun(op, [[copy_state(st)]])
    with P.Pic1dGpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        ref.init_field()
        g.collect_charge()
        g.solve_field()
        for it in range(10):
            for irk in (1, 2):
                ref.push(irk)
                ref.collect_charge()
                ref.solve_field()
                g.push(irk)
                g.collect_charge()
                g.solve_field()
                f = g.get_field()
                assert rel_err(f["chargeden"], ref.rho) < TOL_SUM, (it, irk)
                assert rel_err(f["electric"], ref.E) < TOL_SUM, (it, irk)
                assert np.array_equal(g.get_shape_x(0)[0], ref.o.shape(ref.st[0][0]["x"].copy())[0]), (it, irk)
        out = g.get_markers(0)
    for k in ("x", "v", "w"):
        assert rel_err(out[k], ref.st[0][0][k]) < 1e-12, k


@pytest.mark.parametrize("nx,arith", [(192, 0), (256, 0), (1024, 0), (1024, 1), (4096, 0)])
def test_default_deposit_is_the_reproducible_fixed_point_one(nx, arith):
    """PIC1DP_DEPOSIT_AUTO selects the fixed-point deposit with native 32-bit shared-memory adds (the fastest one on
    B200 at every grid size): two runs of the DEFAULT configuration agree bit for bit in rho, E, x, w, and match the
    oracle; the CAS deposit (fp64 sums in arrival order) agrees with it to the summation tolerance."""
    n = 300_007
    op, gp = make_params(nx=nx, capacity=n, arith_mode=arith)
    st = synth_markers(op, n, seed=27, spread=0.2)
    ref = OracleRun(op, [[copy_state(st)]])
    ref.init_field()
    for _ in range(3):
        ref.step()
    a, b = _run(gp, st, 3), _run(gp, st, 3)
    assert a[2].deposit_mode == P.DEPOSIT_FIXED
    assert np.array_equal(a[0]["chargeden"], b[0]["chargeden"]) and np.array_equal(a[0]["electric"], b[0]["electric"])
    assert np.array_equal(a[1]["w"], b[1]["w"]) and np.array_equal(a[1]["x"], b[1]["x"])
    assert rel_err(a[0]["chargeden"], ref.rho) < TOL_SUM and rel_err(a[0]["electric"], ref.E) < TOL_SUM
    wtol = 1e-12
    for k in ("x", "v", "w"):
        assert rel_err(a[1][k], ref.st[0][0][k]) < wtol, k
    _, gp_cas = make_params(nx=nx, capacity=n, arith_mode=arith, deposit_mode=P.DEPOSIT_SMEM_ATOMIC)
    c = _run(gp_cas, st, 3)
    assert c[2].deposit_mode == P.DEPOSIT_SMEM_ATOMIC
    assert rel_err(c[0]["chargeden"], a[0]["chargeden"]) < TOL_SUM
