"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle on identical marker arrays.

Bars (BASELINE.json north_star): cell indices bit-exact; x, v bit-exact per substep given the same E (no exp on
that path); w within 1e-14 relative (CUDA exp vs glibc exp, <= 1 ulp each); rho, E within 1e-12 * max|.| (fp64
summation order differs: marker order per rank on the CPU, warp/CTA order on the GPU).
"""
import numpy as np
import pytest

import pic1dp_b200 as P
from helpers import OracleRun, copy_state, make_params, rel_err, synth_markers

pytestmark = pytest.mark.gpu

TOL_SUM = 1e-12   # rho, E, modes: relative to max|.|
TOL_W = 1e-14     # w after one substep, relative to max|w|
DEPOSITS = [P.DEPOSIT_SMEM_ATOMIC, P.DEPOSIT_GLOBAL_RED, P.DEPOSIT_WARP_PRIVATE, P.DEPOSIT_FIXED]


def _gpu(gp):
    return P.Pic1dGpu(gp)


def test_operators_bit_exact():
    """cos / -sin tables and 1/k (src/pic1dp_field.F90:158-210) are bit-identical to the restatement."""
    from oracle import oracle as O
    for nx, modes in ((192, [1]), (256, [1, 2, 5]), (1000, [3])):
        op, gp = make_params(nx=nx, nmode=len(modes), modes=modes, capacity=16)
        o = O.Oracle(op)
        with _gpu(gp) as g:
            Fre, Fim, ginv = g.get_operators()
        assert np.array_equal(Fre.ravel(), o.F_re) and np.array_equal(Fim.ravel(), o.F_im)
        assert np.array_equal(ginv, o.grad_inv)


@pytest.mark.parametrize("nx,modes", [(192, [1]), (256, [1, 2, 3, 7]), (4096, [1]), (8192, [1, 4])])
def test_field_solve_sequential_bit_exact(nx, modes):
    """Given the same rho, the sequential-order solve reproduces E, mode_re, mode_im bit for bit."""
    from oracle import oracle as O
    op, gp = make_params(nx=nx, nmode=len(modes), modes=modes, capacity=16, field_mode=P.FIELD_SEQUENTIAL)
    rho = np.random.default_rng(1).standard_normal(nx) * 1e-4
    E, mre, mim = O.Oracle(op).field_solve(rho)
    with _gpu(gp) as g:
        g.set_field(chargeden=rho)
        g.solve_field()
        f = g.get_field()
    assert np.array_equal(f["electric"], E)
    assert np.array_equal(f["mode_re"], mre) and np.array_equal(f["mode_im"], mim)


@pytest.mark.parametrize("nx,modes", [(192, [1]), (1024, [1, 2, 3]), (8192, [1])])
def test_field_solve_tree_tolerance(nx, modes):
    from oracle import oracle as O
    op, gp = make_params(nx=nx, nmode=len(modes), modes=modes, capacity=16, field_mode=P.FIELD_TREE)
    rho = np.random.default_rng(2).standard_normal(nx) * 1e-4
    E, mre, mim = O.Oracle(op).field_solve(rho)
    with _gpu(gp) as g:
        g.set_field(chargeden=rho)
        g.solve_field()
        f = g.get_field()
        f2 = None
        g.solve_field()
        f2 = g.get_field()
    assert rel_err(f["electric"], E) < TOL_SUM
    scale = max(np.abs(mre).max(), np.abs(mim).max())
    assert rel_err(f["mode_re"], mre, scale) < TOL_SUM and rel_err(f["mode_im"], mim, scale) < TOL_SUM
    assert np.array_equal(f["electric"], f2["electric"])  # fixed-shape tree: run-to-run identical


def test_field_test_analytic():
    """field_test (src/pic1dp_field.F90:276-309): rho = cos(2 pi j / nx) -> E = sin(2 pi j / nx) / k1."""
    op, gp = make_params(capacity=16)
    j = np.arange(gp.nx)
    with _gpu(gp) as g:
        g.set_field(chargeden=np.cos(2.0 * np.pi * j / gp.nx))
        g.solve_field()
        E = g.get_field()["electric"]
    assert np.max(np.abs(E - np.sin(2.0 * np.pi * j / gp.nx) / 0.36)) < 1e-13
    m = P.Pic1dpModules(gp)   # the same through the module mirror's field_test
    m.field_init()
    E2 = m.field_test()
    m.field_final()
    assert np.array_equal(E, E2)


@pytest.mark.parametrize("shape", [4, 1])
def test_cell_index_and_weights_bit_exact(shape):
    """ix, left and right weights for every marker, including wrapped / edge coordinates."""
    from oracle import oracle as O
    op, gp = make_params(nx=256, iptclshape=shape, capacity=300000)
    st = synth_markers(op, 250001, seed=3)
    lx = op.lx
    edge = np.array([0.0, -0.0, lx, np.nextafter(lx, 0), lx / 256, np.nextafter(lx / 256, 0), lx * 0.5,
                     np.nextafter(lx, 0) * 0.5, 5e-324, lx * (255.0 / 256.0)])
    st["x"][:edge.size] = edge
    o = O.Oracle(op)
    xo = st["x"].copy()
    ix, sl, sr = o.shape(xo, right_frac=(shape <= 2))  # wraps (no-op here: all within [0, lx])
    with _gpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        gix, gsl, gsr = g.get_shape_x(0)
    ok = ix < op.nx  # x == lx: reference is out of bounds there; both sides define it as cell 0, s = 1
    assert np.array_equal(gix[ok], ix[ok])
    assert np.array_equal(gsl[ok], sl[ok]) and np.array_equal(gsr[ok], sr[ok])
    assert np.all(gix[~ok] == 0) and np.all(gsl[~ok] == 1.0)


@pytest.mark.parametrize("dep", DEPOSITS)
@pytest.mark.parametrize("nx", [192, 1024])
def test_collect_charge_standalone(dep, nx):
    """interaction_collect_charge on freshly loaded markers: wrapped x bit-exact, rho within tolerance."""
    op, gp = make_params(nx=nx, capacity=200000, deposit_mode=dep)
    st = synth_markers(op, 150001, seed=4, spread=1.5)
    ref = OracleRun(op, [[copy_state(st)]])
    ref.collect_charge()
    with _gpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        g.collect_charge()
        rho = g.get_field()["chargeden"]
        x = g.get_markers(0, want=("x",))["x"]
        assert g.counters().deposit_mode == dep
        noob = g.counters().oob_markers
    assert np.array_equal(x, ref.st[0][0]["x"])
    assert rel_err(rho, ref.rho) < TOL_SUM
    assert noob == ref.noob


@pytest.mark.parametrize("dist", [0, 1, 2, 3])
@pytest.mark.parametrize("fuse", [0, 1])
def test_push_substeps_against_oracle(dist, fuse):
    """Both RK substeps with a prescribed E: x, v bit-exact; w bit-exact without exp (dist 0, 1), 1e-14 otherwise.
    fuse=0 keeps the reference's side effects (x left unwrapped); fuse=1 is compared after the oracle's wrap."""
    op, gp = make_params(nx=256, iptcldist=dist, capacity=100001, fuse=fuse, temperature=[1.3], temperature2=[0.7],
                         mass=[1.0] if dist != 2 else [2.0])
    n = 100001
    st = synth_markers(op, n, seed=5 + dist)
    if dist == 1:
        st["v"][np.abs(st["v"]) < 1e-3] = 0.5  # 2/v
    E = 1e-3 * np.sin(2 * np.pi * np.arange(op.nx) / op.nx + 0.3)
    ref = OracleRun(op, [[copy_state(st)]])
    ref.E = E.copy()
    with _gpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        g.set_field(electric=E)
        for irk in (1, 2):
            ref.push(irk)
            g.push(irk)
            out = g.get_markers(0)
            r = ref.st[0][0]
            if fuse:
                # the fused kernel has already applied the wrap of the following collect_charge
                xr = r["x"].copy()
                ref.o.shape(xr)
                assert np.array_equal(out["x"], xr)
                r["x"][:] = xr  # the reference wraps here too (collect_charge) before the next push
            else:
                assert np.array_equal(out["x"], r["x"])
                ref.collect_charge()     # wraps x in place on the oracle side
                g.collect_charge()       # and on the GPU side
                ref.E = E.copy()
                g.set_field(electric=E)
            assert np.array_equal(out["v"], r["v"])
            if dist in (0, 1):
                assert np.array_equal(out["w"], r["w"])
            else:
                assert rel_err(out["w"], r["w"]) < TOL_W
                r["w"][:] = out["w"]  # continue from identical state so irk=2 isolates one substep
            assert np.array_equal(out["p"], r["p"])


@pytest.mark.parametrize("dep", DEPOSITS)
@pytest.mark.parametrize("fuse", [0, 1])
def test_ten_steps_default_physics(dep, fuse):
    """Config C2 physics at test size: default bump-on-tail, nx=256, the reference driver sequence for 10 steps.
    rho, E, modes within 1e-12 of max; x, v, w within 1e-12 relative; cell indices identical at every substep."""
    op, gp = make_params(nx=256, capacity=200000, deposit_mode=dep, fuse=fuse)
    st = synth_markers(op, 200000, seed=11)
    ref = OracleRun(op, [[copy_state(st)]])
    m = P.Pic1dpModules(gp)
    m.particle_init()
    m.field_init()
    m.particle_set(0, st["x"], st["v"], st["p"], st["w"])
    ref.init_field()
    m.interaction_collect_charge()
    m.field_solve_electric()
    assert rel_err(m.field_chargeden, ref.rho) < TOL_SUM
    assert rel_err(m.field_electric, ref.E) < TOL_SUM
    for it in range(10):
        for irk in (1, 2):
            ref.push(irk)
            ref.collect_charge()
            ref.solve_field()
            m.global_irk = irk
            m.interaction_push_particle()
            m.interaction_collect_charge()
            m.field_solve_electric()
            assert rel_err(m.field_chargeden, ref.rho) < TOL_SUM, (it, irk)
            assert rel_err(m.field_electric, ref.E) < TOL_SUM, (it, irk)
            gix, _, _ = m.gpu.get_shape_x(0)
            rix, _, _ = ref.o.shape(ref.st[0][0]["x"].copy())
            assert np.array_equal(gix, rix), (it, irk)
    out = m.particle_get(0)
    r = ref.st[0][0]
    for k in ("x", "v", "w"):
        assert rel_err(out[k], r[k]) < 1e-12, k
    scale = max(np.abs(ref.mode_re).max(), np.abs(ref.mode_im).max())
    assert rel_err(m.field_mode_re, ref.mode_re, scale) < TOL_SUM
    assert rel_err(m.field_mode_im, ref.mode_im, scale) < TOL_SUM
    assert m.gpu.counters().oob_markers == ref.noob == 0
    m.particle_final()


def test_step_equals_individual_calls():
    op, gp = make_params(nx=192, capacity=50000, deposit_mode=P.DEPOSIT_WARP_PRIVATE)
    st = synth_markers(op, 50000, seed=12)
    outs = []
    for use_step in (False, True):
        with _gpu(gp) as g:
            g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
            g.collect_charge()
            g.solve_field()
            if use_step:
                g.step(3)
            else:
                for _ in range(3):
                    for irk in (1, 2):
                        g.push(irk)
                        g.collect_charge()
                        g.solve_field()
            outs.append((g.get_markers(0), g.get_field()))
    for k in ("x", "v", "w"):
        assert np.array_equal(outs[0][0][k], outs[1][0][k])
    assert np.array_equal(outs[0][1]["electric"], outs[1][1]["electric"])


def test_warp_private_deposit_is_bitwise_deterministic():
    op, gp = make_params(nx=512, capacity=300000, deposit_mode=P.DEPOSIT_WARP_PRIVATE)
    st = synth_markers(op, 300000, seed=13)
    res = []
    for _ in range(3):
        with _gpu(gp) as g:
            g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
            g.collect_charge()
            g.solve_field()
            g.step(4)
            res.append((g.get_field(), g.get_markers(0)))
    for f, mk in res[1:]:
        assert np.array_equal(f["chargeden"], res[0][0]["chargeden"])
        assert np.array_equal(f["electric"], res[0][0]["electric"])
        assert np.array_equal(mk["w"], res[0][1]["w"])


@pytest.mark.parametrize("case", ["fullf", "linear", "two_species", "matrix_shape", "multi_mode", "landau_4096",
                                  "pow2_nonunit", "nonpow2"])
def test_model_variants_three_steps(case):
    kw = dict(nx=256, capacity=60000)
    nsp = 1
    if case == "fullf":
        kw.update(deltaf=0, iptcldist=0, density=[1.0], v0=[0.0])
    elif case == "linear":
        kw.update(linear=1)
    elif case == "two_species":
        nsp = 2
        kw.update(nspecies=2, charge=[-1.0, 1.0], mass=[1.0, 4.0], temperature=[1.0, 0.5], temperature2=[1.0, 0.5],
                  density=[0.9, 1.0], v0=[5.0, 0.0], iptcldist=3)
    elif case == "matrix_shape":
        kw.update(iptclshape=2)
    elif case == "multi_mode":
        kw.update(nmode=3, modes=[1, 2, 3])
    elif case == "pow2_nonunit":   # every constant divisor a power of two but not 1: exact reciprocal multiplies
        kw.update(temperature=[2.0], mass=[0.5], temperature2=[0.5])
    elif case == "nonpow2":        # true divisions by T/m, sqrt(T/m), ...
        kw.update(temperature=[1.3], mass=[0.9], temperature2=[0.7])
    elif case == "landau_4096":
        kw.update(nx=4096, iptcldist=0, density=[1.0], v0=[0.0], lx=4 * np.pi)
    op, gp = make_params(**kw)
    sts = [synth_markers(op, 60000 - 7 * s, seed=20 + s, isp=s) for s in range(nsp)]
    ref = OracleRun(op, [[copy_state(s)] for s in sts])
    ref.init_field()
    with _gpu(gp) as g:
        for s, st in enumerate(sts):
            g.set_markers(s, st["x"], st["v"], st["p"], st["w"])
        if gp.iptclshape < 4:
            g.compute_shape_x()
        g.collect_charge()
        g.solve_field()
        f = g.get_field()
        assert rel_err(f["chargeden"], ref.rho) < TOL_SUM
        for _ in range(3):
            ref.step()
        g.step(3)
        f = g.get_field()
        assert rel_err(f["chargeden"], ref.rho) < TOL_SUM
        assert rel_err(f["electric"], ref.E) < TOL_SUM
        for s in range(nsp):
            out = g.get_markers(s)
            for k in ("x", "v", "w", "p"):
                assert rel_err(out[k], ref.st[s][0][k]) < 1e-12, (case, s, k)


def test_ragged_and_empty_inputs():
    """np = 0, 1, odd, one full tile +- 1; capacity overflow and call-order errors."""
    op, gp = make_params(nx=192, capacity=5000)
    with _gpu(gp) as g:
        with pytest.raises(P.Pic1dpError) as e:
            g.push(1)
        assert e.value.code == 5  # ESTATE: markers not set
        big = synth_markers(op, 5001, seed=1)
        with pytest.raises(P.Pic1dpError) as e:
            g.set_markers(0, big["x"], big["v"], big["p"], big["w"])
        assert e.value.code == 6  # ECAPACITY
    for n in (0, 1, 2, 3, 1023, 1024, 1025, 4999):
        st = synth_markers(op, n, seed=30 + n, spread=0.7)
        ref = OracleRun(op, [[copy_state(st)]])
        ref.init_field()
        ref.step()
        with _gpu(gp) as g:
            g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
            g.collect_charge()
            g.solve_field()
            g.step(1)
            f = g.get_field()
            out = g.get_markers(0)
        assert out["x"].size == n
        scale = max(np.abs(ref.rho).max(), 1e-300)
        assert rel_err(f["chargeden"], ref.rho, scale) < TOL_SUM, n
        for k in ("x", "v", "w"):
            assert rel_err(out[k], ref.st[0][0][k]) < 1e-12, (n, k)


def test_far_out_of_box_coordinates_wrap_like_fmod():
    op, gp = make_params(nx=192, capacity=4096)
    st = synth_markers(op, 4096, seed=40)
    lx = op.lx
    st["x"][:8] = [-3.7 * lx, 12.25 * lx, -lx, 2 * lx, -1e-300, 1e6, -1e6, np.nextafter(-lx, 0)]
    ref = OracleRun(op, [[copy_state(st)]])
    ref.collect_charge()
    with _gpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        g.collect_charge()
        x = g.get_markers(0, want=("x",))["x"]
        rho = g.get_field()["chargeden"]
        noob = g.counters().oob_markers
    assert np.array_equal(x, ref.st[0][0]["x"])
    assert rel_err(rho, ref.rho) < TOL_SUM
    assert noob == ref.noob


def test_charge_conservation_and_linearity_at_scale():
    """Size-independent properties at C2 size (1e7 markers, nx=256): sum_j rho_j * lx/nx == Z * sum w (weights
    sum to 1), and deposit(w1) + deposit(w2) == deposit(w1 + w2), both to summation-order tolerance."""
    n = 10_000_000
    op, gp = make_params(nx=256, capacity=n)
    st = synth_markers(op, n, seed=50)
    w2 = np.cos(st["x"]) * 1e-12
    with _gpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        g.collect_charge()
        r1 = g.get_field()["chargeden"]
        g.set_markers(0, st["x"], st["v"], st["p"], w2)
        g.collect_charge()
        r2 = g.get_field()["chargeden"]
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"] + w2)
        g.collect_charge()
        r12 = g.get_field()["chargeden"]
    tot = np.sum(r1) * op.lx / op.nx
    expect = op.charge[0] * np.sum(st["w"])
    sabs = np.sum(np.abs(st["w"]))
    assert abs(tot - expect) < 1e-12 * sabs
    assert rel_err(r1 + r2, r12) < TOL_SUM


@pytest.mark.parametrize("load_path", [P._capi.LOAD_TMA, P._capi.LOAD_CPASYNC])
@pytest.mark.parametrize("dep", DEPOSITS)
@pytest.mark.parametrize("case", ["unit", "pow2", "nonpow2", "maxwell"])
def test_staged_load_paths_against_oracle(dep, case, load_path):
    """load_path = TMA (cp.async.bulk tiles through a shared-memory ring) and CPASYNC (per-thread cp.async into
    thread-private ring slots): one substep with prescribed E is bit-exact in x, v; five full steps stay within the
    summation-order tolerance; ragged sizes cover partial tiles."""
    kw = dict(nx=256, capacity=1000003, deposit_mode=dep, load_path=load_path)
    if case == "pow2":
        kw.update(temperature=[2.0], mass=[0.5], temperature2=[0.5])
    elif case == "nonpow2":
        kw.update(temperature=[1.3], mass=[0.9], temperature2=[0.7])
    elif case == "maxwell":
        kw.update(iptcldist=0, density=[1.0], v0=[0.3])
    op, gp = make_params(**kw)
    try:
        _gpu(gp).close()
    except P.Pic1dpError as e:
        assert e.code == 8  # PIC1DP_EUNSUPPORTED: the ring does not reach the direct kernel's residency here
        pytest.skip("shared-memory ring not available for this deposit mode / nx")
    for n in (1000003, 70001, 512, 513, 1, 1023):  # 1000003: several tile steps per CTA, so the ring wraps
        st = synth_markers(op, n, seed=80 + n % 7)
        E = 1e-3 * np.sin(2 * np.pi * np.arange(op.nx) / op.nx + 0.1)
        ref = OracleRun(op, [[copy_state(st)]])
        ref.E = E.copy()
        with _gpu(gp) as g:
            g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
            g.set_field(electric=E)
            ref.push(1)
            g.push(1)
            out = g.get_markers(0)
            xr = ref.st[0][0]["x"].copy()
            ref.o.shape(xr)
            assert np.array_equal(out["x"], xr) and np.array_equal(out["v"], ref.st[0][0]["v"]), n
            assert rel_err(out["w"], ref.st[0][0]["w"]) < TOL_W
        ref = OracleRun(op, [[copy_state(st)]])
        ref.init_field()
        for _ in range(5):
            ref.step()
        with _gpu(gp) as g:
            g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
            g.collect_charge()
            g.solve_field()
            g.step(5)
            f = g.get_field()
            out = g.get_markers(0)
        scale = max(np.abs(ref.rho).max(), 1e-300)
        assert rel_err(f["chargeden"], ref.rho, scale) < TOL_SUM, n
        for k in ("x", "v", "w"):
            assert rel_err(out[k], ref.st[0][0][k]) < 1e-12, (n, k)


@pytest.mark.parametrize("dep", DEPOSITS)
@pytest.mark.parametrize("nx,nmode", [(2, 1), (3, 1), (193, 2), (1001, 3), (64, 32)])
def test_odd_and_extreme_grid_sizes(dep, nx, nmode):
    """nx = 2 (both neighbours are the only other cell), odd nx (16-byte alignment of the pair grid after the E copy),
    many kept modes."""
    modes = list(range(1, nmode + 1)) if nmode < nx else [1]
    op, gp = make_params(nx=nx, nmode=len(modes), modes=modes, capacity=30011, deposit_mode=dep)
    st = synth_markers(op, 30011, seed=90 + nx, spread=0.6)
    ref = OracleRun(op, [[copy_state(st)]])
    ref.init_field()
    for _ in range(3):
        ref.step()
    with _gpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        g.collect_charge()
        g.solve_field()
        g.step(3)
        f = g.get_field()
        out = g.get_markers(0)
    assert rel_err(f["chargeden"], ref.rho) < TOL_SUM
    assert rel_err(f["electric"], ref.E, max(np.abs(ref.E).max(), 1e-300)) < 1e-11
    for k in ("x", "v", "w"):
        assert rel_err(out[k], ref.st[0][0][k]) < 1e-12, k


def test_max_species_and_modes_limits():
    op, gp = make_params(nx=128, nspecies=4, charge=[-1.0, 1.0, -1.0, 2.0], mass=[1.0, 4.0, 1.0, 8.0],
                         temperature=[1.0, 0.5, 2.0, 1.0], temperature2=[1.0, 0.5, 1.0, 1.0],
                         density=[0.9, 1.0, 0.5, 0.25], v0=[5.0, 0.0, 3.0, 1.0], nmode=4, modes=[1, 2, 3, 4], capacity=20000)
    sts = [synth_markers(op, 20000 - 3 * s, seed=100 + s, isp=s) for s in range(4)]
    ref = OracleRun(op, [[copy_state(s)] for s in sts])
    ref.init_field()
    ref.step()
    ref.step()
    with _gpu(gp) as g:
        for s, st in enumerate(sts):
            g.set_markers(s, st["x"], st["v"], st["p"], st["w"])
        g.collect_charge()
        g.solve_field()
        g.step(2)
        f = g.get_field()
        assert rel_err(f["chargeden"], ref.rho) < TOL_SUM and rel_err(f["electric"], ref.E) < TOL_SUM
        for s in range(4):
            out = g.get_markers(s)
            for k in ("x", "v", "w"):
                assert rel_err(out[k], ref.st[s][0][k]) < 1e-12, (s, k)


@pytest.mark.parametrize("dist", [0, 1, 2, 3])
@pytest.mark.parametrize("linear", [0, 1])
def test_device_side_particle_load_matches_oracle_loader(dist, linear):
    """pic1dp_gpu_load_markers given the two multirand streams (SuperKISS64, constant seeds, rank 1) against the oracle's
    restated particle_load: x, v bit-exact; p, w within a few ulp (device exp / sin / cos vs glibc)."""
    from oracle import oracle as O
    n, ntot = 100003, 400012
    op, gp = make_params(nx=192, iptcldist=dist, linear=linear, capacity=n, temperature=[1.3], temperature2=[0.8],
                         mass=[1.1], density=[0.85], v0=[2.5], init_nmode=2, init_mode=[1, 3],
                         init_mode_cos=[2e-6, 0.0], init_mode_sin=[1e-5, 3e-6])
    o = O.Oracle(op)
    x, v, p, w = o.particle_load(0, 3, 1, 5, n, ntot)
    rng = O.MultiRand()
    rng.init_const(3, 1, 5)
    rand_v = rng.real_array(n)   # drawn first (src/pic1dp_particle.F90:180), then x (:222)
    rand_x = rng.real_array(n)
    with _gpu(gp) as g:
        g.load_markers(0, rand_v, rand_x, ntot, v_max=8.0, init_mode=(1, 3), init_cos=(2e-6, 0.0), init_sin=(1e-5, 3e-6))
        out = g.get_markers(0)
        assert g.counters().h2d_bytes == 16 * n
    assert np.array_equal(out["x"], x) and np.array_equal(out["v"], v)
    assert rel_err(out["p"], p) < 1e-14
    assert rel_err(out["w"], w) < 1e-13
    if dist == 3 and linear == 0:  # the same through the module mirror's particle_load
        m = P.Pic1dpModules(gp)
        m.particle_init()
        m.particle_load(0, rand_v, rand_x, ntot, 8.0, (1, 3), (2e-6, 0.0), (1e-5, 3e-6))
        out2 = m.particle_get(0)
        m.particle_final()
        assert m.particle_np[0] == n and all(np.array_equal(out[k], out2[k]) for k in ("x", "v", "p", "w"))


@pytest.mark.parametrize("linear", [0, 1])
def test_device_side_particle_load_maxwellian_markers(linear):
    """input_imarker = 1 (src/pic1dp_particle.F90:172-178): v from the Gaussian stream, constant p -- against the oracle's
    loader; x, v bit-exact, p and w within a few ulp (device sin / cos vs glibc)."""
    from oracle import oracle as O
    n, ntot = 100003, 400012
    op, gp = make_params(nx=192, iptcldist=0, linear=linear, capacity=n, temperature=[1.3], mass=[1.1], density=[0.85],
                         v0=[2.5], imarker=1, init_nmode=2, init_mode=[1, 3], init_mode_cos=[2e-6, 0.0],
                         init_mode_sin=[1e-5, 3e-6])
    x, v, p, w = O.Oracle(op).particle_load(0, 3, 1, 5, n, ntot)
    rng = O.MultiRand()
    rng.init_const(3, 1, 5)
    gauss_v = rng.gaussian_array(n)   # multirand_gaussian_array(pv) :174, then the uniform x stream :222
    rand_x = rng.real_array(n)
    with _gpu(gp) as g:
        g.load_markers_maxwellian(0, gauss_v, rand_x, ntot, init_mode=(1, 3), init_cos=(2e-6, 0.0), init_sin=(1e-5, 3e-6))
        out = g.get_markers(0)
    assert np.array_equal(out["x"], x) and np.array_equal(out["v"], v)
    assert rel_err(out["p"], p) < 1e-14 and rel_err(out["w"], w) < 1e-13
    assert abs(np.mean(v) - 2.5) < 0.02 and abs(np.std(v) / np.sqrt(1.3 / 1.1) - 1.0) < 0.01
    _, gp3 = make_params(nx=192, iptcldist=3, capacity=n)   # input_init rejects imarker = 1 with iptcldist >= 1
    with _gpu(gp3) as g:
        with pytest.raises(P.Pic1dpError) as e:
            g.load_markers_maxwellian(0, gauss_v, rand_x, ntot)
        assert e.value.code == 1


def test_randomized_configurations_two_steps():
    """24 pseudo-random parameter sets (fixed seed): equilibrium, delta-f / full-f, linear, weight rounding, deposit
    mode, fuse, species constants, grid size, mode set, ragged marker counts -- two full steps against the oracle."""
    rng = np.random.default_rng(2024)
    for trial in range(24):
        dist = int(rng.integers(0, 4))
        deltaf = int(rng.integers(0, 2)) if dist in (0, 1) else 1
        linear = int(rng.integers(0, 2)) if deltaf else 0
        nx = int(rng.choice([2, 7, 64, 192, 333, 1024, 2048]))
        nmode = int(rng.integers(1, 4)) if nx >= 16 else 1
        modes = sorted(rng.choice(np.arange(1, max(2, min(9, nx // 2 + 1))), size=nmode, replace=False).tolist())
        dep = int(rng.choice([P.DEPOSIT_AUTO, P.DEPOSIT_SMEM_ATOMIC, P.DEPOSIT_GLOBAL_RED, P.DEPOSIT_WARP_PRIVATE, P.DEPOSIT_FIXED]))
        n = int(rng.choice([1, 33, 1000, 4097, 20001]))
        kw = dict(nx=nx, nmode=nmode, modes=modes, iptcldist=dist, deltaf=deltaf, linear=linear,
                  iptclshape=int(rng.choice([1, 2, 3, 4])), deposit_mode=dep, fuse=int(rng.integers(0, 2)),
                  temperature=[float(rng.choice([1.0, 2.0, 1.3]))], temperature2=[float(rng.choice([1.0, 0.5, 0.7]))],
                  mass=[float(rng.choice([1.0, 0.5, 1.7]))], density=[float(rng.choice([0.9, 1.0, 0.6]))],
                  v0=[float(rng.choice([5.0, 0.0, 2.0]))], charge=[float(rng.choice([-1.0, 1.0, 2.0]))],
                  dt=float(rng.choice([0.05, 0.1])), capacity=max(n, 1))
        op, gp = make_params(**kw)
        st = synth_markers(op, n, seed=1000 + trial, spread=float(rng.choice([0.0, 0.4])))
        if dist == 1:
            st["v"][np.abs(st["v"]) < 1e-2] = 0.7
        ref = OracleRun(op, [[copy_state(st)]])
        ref.init_field()
        ref.step()
        ref.step()
        with _gpu(gp) as g:
            g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
            if gp.iptclshape < 4:
                g.compute_shape_x()
            g.collect_charge()
            g.solve_field()
            g.step(2)
            f = g.get_field()
            out = g.get_markers(0)
            noob = g.counters().oob_markers
        tag = (trial, kw)
        scale = max(np.abs(ref.rho).max(), 1e-300)
        assert rel_err(f["chargeden"], ref.rho, scale) < 1e-11, tag
        assert rel_err(f["electric"], ref.E, max(np.abs(ref.E).max(), 1e-300)) < 1e-10, tag
        for k in ("x", "v", "w", "p"):
            assert rel_err(out[k], ref.st[0][0][k]) < 1e-11, (k, tag)
        assert noob == ref.noob, tag


def test_committed_hotpath_fixture():
    """The GPU against the committed oracle-generated fixture tests/golden/hotpath_tiny.json (no oracle call)."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hotpath_tiny.json")))
    fh = lambda a: np.array([float.fromhex(t) for t in a])
    _, gp = make_params(nx=g["nx"], capacity=g["n"], field_mode=P.FIELD_SEQUENTIAL)
    with _gpu(gp) as gpu:
        gpu.set_markers(0, *(fh(g["init"][k]) for k in ("x", "v", "p", "w")))
        gpu.collect_charge()
        gpu.solve_field()
        f0 = gpu.get_field()
        gpu.step(g["steps"])
        f = gpu.get_field()
        out = gpu.get_markers(0)
    assert rel_err(f0["chargeden"], fh(g["after"]["rho0"])) < TOL_SUM and rel_err(f0["electric"], fh(g["after"]["E0"])) < TOL_SUM
    assert rel_err(f["chargeden"], fh(g["after"]["rho"])) < TOL_SUM and rel_err(f["electric"], fh(g["after"]["E"])) < TOL_SUM
    for k in ("x", "v", "w"):
        assert rel_err(out[k], fh(g["after"][k])) < 1e-12, k
