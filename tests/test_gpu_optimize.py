"""Marker optimisation through the C ABI on the GPU (/root/reference/src/pic1dp_particle.F90:356-813):
the device reduction particle_compute_dist_pertb_abs_v against the oracle, merge / remove / split on device-resident
markers against the oracle bit for bit, and the reference time loop with optimisation events (src/pic1dp.F90:78-93)
against the oracle's replay of the same loop."""
import numpy as np
import pytest

import pic1dp_b200 as P
from helpers import OracleRun, copy_state, make_params, rel_err, synth_markers
from oracle import oracle as O

pytestmark = pytest.mark.gpu

VMAX = 8.0


def _developed(op, n, seed):
    st = synth_markers(op, n, seed=seed, spread=0.2)
    st["w"] = st["w"] * (1.0 + 50.0 * np.exp(-(st["v"] - 3.2) ** 2)) * np.sign(np.sin(7.0 * st["x"]) + 0.3)
    return st


@pytest.mark.parametrize("nv,n", [(128, 300001), (64, 77777), (1000, 50000), (2, 1000), (128, 31)])
def test_dist_pertb_abs_v_matches_oracle_and_is_deterministic(nv, n):
    op, gp = make_params(nx=192, capacity=n)
    st = _developed(op, n, seed=nv)
    st["v"][:6] = [np.nextafter(VMAX, 0.0), -np.nextafter(VMAX, 0.0), VMAX, -VMAX, 9.5, -0.0][:min(6, n)]
    ref = O.Oracle(op).dist_pertb_abs_v([st["v"]], [st["w"]], nv, VMAX)
    with P.Pic1dGpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        d1 = g.compute_dist_pertb_abs_v(nv, VMAX)
        d2 = g.compute_dist_pertb_abs_v(nv, VMAX)
    assert d1.shape == (1, nv)
    assert rel_err(d1[0], ref) < 1e-12
    assert np.array_equal(d1, d2), "fixed summation order: bitwise run-to-run"


def test_dist_two_species():
    op, gp = make_params(nx=64, capacity=40000, nspecies=2, charge=[-1.0, 1.0], mass=[1.0, 4.0], temperature=[1.0, 0.5],
                         temperature2=[1.0, 0.5], density=[0.9, 1.0], v0=[5.0, 0.0])
    sts = [_developed(op, 40000 - 7 * s, seed=5 + s) for s in range(2)]
    with P.Pic1dGpu(gp) as g:
        for s, st in enumerate(sts):
            g.set_markers(s, st["x"], st["v"], st["p"], st["w"])
        d = g.compute_dist_pertb_abs_v(128, VMAX)
    for s, st in enumerate(sts):
        assert rel_err(d[s], O.Oracle(op).dist_pertb_abs_v([st["v"]], [st["w"]], 128, VMAX)) < 1e-12


def test_optimise_needs_dist_first():
    op, gp = make_params(nx=64, capacity=1000)
    st = _developed(op, 1000, 1)
    with P.Pic1dGpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        with pytest.raises(P.Pic1dpError) as e:
            g.particle_merge(0.1)
        assert e.value.code == 5  # PIC1DP_ESTATE


@pytest.mark.parametrize("fuse", [1, 0])
def test_merge_remove_split_on_device_markers_match_oracle_bit_for_bit(fuse):
    n, cap = 200000, 260000
    op, gp = make_params(nx=64, capacity=cap, fuse=fuse)
    st = _developed(op, n, seed=21)
    orc = O.Oracle(op)
    r_ref, r_gpu = O.MultiRand(), O.MultiRand()
    r_ref.init_const(3, 0, 5)
    r_gpu.init_const(3, 0, 5)
    with P.Pic1dGpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        g.collect_charge()
        g.solve_field()
        g.step(2)  # the markers the optimiser sees come out of the device push

        def check(op_name, run_gpu, run_ref):
            dist = g.compute_dist_pertb_abs_v(128, VMAX)[0]
            pre = g.get_markers(0)
            n0 = pre["x"].size
            ref = {k: np.concatenate([a, np.full(cap - n0, np.nan)]) for k, a in pre.items()}
            n_ref = run_ref(ref, n0, dist)
            n_gpu = run_gpu()[0]
            assert n_gpu == n_ref, op_name
            post = g.get_markers(0)
            for k in ("x", "v", "p", "w"):
                assert post[k].size == n_ref and np.array_equal(post[k], ref[k][:n_ref]), (op_name, k)
            return n0, n_ref

        n0, n1 = check("merge", lambda: g.particle_merge(0.3), lambda s, m, d: orc.particle_merge(s, m, d, 0.3, VMAX))
        assert n1 < n0
        n1b, n2 = check("remove", lambda: g.particle_remove(0.0, 2, 0.9, r_gpu.real64),
                        lambda s, m, d: orc.particle_remove(s, m, d, 0.0, 2, 0.9, r_ref, VMAX))
        assert n1b == n1 and n2 < n1
        _, n3 = check("split", lambda: g.particle_split(0.5, 5, 0.1, r_gpu.gaussian_array),
                      lambda s, m, d: orc.particle_split(s, m, d, 0.5, 5, 0.1, r_ref, VMAX))
        assert n2 < n3 <= cap
        assert r_ref.int64() == r_gpu.int64()
        # the path keeps running on the new marker set
        g.collect_charge()
        g.solve_field()
        g.step(1)
        assert g.get_markers(0)["x"].size == n3 and np.isfinite(g.field_energy())


@pytest.mark.parametrize("fuse,dep", [(1, P.DEPOSIT_AUTO), (0, P.DEPOSIT_WARP_PRIVATE), (1, P.DEPOSIT_GLOBAL_RED)])
def test_time_loop_with_optimisation_events_matches_oracle(fuse, dep):
    """src/pic1dp.F90:78-93 with particle_optimize between push and collect_charge at irk == 2: a merge at t = 0.1,
    a removal at t = 0.2 and a split at t = 0.3 (dt = 0.05), GPU modules vs the oracle's replay."""
    n, cap, nsteps = 150000, 200000, 8
    op, gp = make_params(nx=128, capacity=cap, fuse=fuse, deposit_mode=dep)
    st = _developed(op, n, seed=33)
    sched = dict(tmerge=[0.1], thshmerge=[0.3], tremove=[0.2], thshremove=[0.0], typeremove=2, remove_frac=0.9,
                 tsplit=[0.3], thshsplit=[0.5], split_ngroup=5, split_dv_sig_frac=0.1, nv=128, v_max=VMAX)
    # ---- oracle replay ----
    ref = OracleRun(op, [[{k: np.concatenate([a, np.zeros(cap - n)]) for k, a in copy_state(st).items()}]])
    rs = ref.st[0][0]
    r_ref = O.MultiRand()
    r_ref.init_const(3, 0, 5)
    npar = n

    def view(m):  # the first m markers as views the oracle updates in place
        return [[{k: a[:m] for k, a in rs.items()}]]

    full = ref.st
    ref.st = view(npar)
    ref.init_field()
    imerge = iremove = isplit = 0
    t, np_trace, e_ref = 0.0, [], []
    for it in range(nsteps):
        for irk in (1, 2):
            ref.st = view(npar)
            ref.push(irk)
            if irk == 2:
                if imerge < 1 and t + op.dt >= sched["tmerge"][0]:
                    d = ref.o.dist_pertb_abs_v([rs["v"][:npar]], [rs["w"][:npar]], 128, VMAX)
                    npar = ref.o.particle_merge(rs, npar, d, 0.3, VMAX)
                    imerge += 1
                if iremove < 1 and t + op.dt >= sched["tremove"][0]:
                    d = ref.o.dist_pertb_abs_v([rs["v"][:npar]], [rs["w"][:npar]], 128, VMAX)
                    npar = ref.o.particle_remove(rs, npar, d, 0.0, 2, 0.9, r_ref, VMAX)
                    iremove += 1
                if isplit < 1 and t + op.dt >= sched["tsplit"][0]:
                    d = ref.o.dist_pertb_abs_v([rs["v"][:npar]], [rs["w"][:npar]], 128, VMAX)
                    npar = ref.o.particle_split(rs, npar, d, 0.5, 5, 0.1, r_ref, VMAX)
                    isplit += 1
            ref.st = view(npar)
            ref.collect_charge()
            ref.solve_field()
        t += op.dt
        np_trace.append(npar)
        e_ref.append(ref.o.field_energy(ref.E))
    assert (imerge, iremove, isplit) == (1, 1, 1) and len(set(np_trace)) >= 4
    ref.st = full
    # ---- GPU modules ----
    r_gpu = O.MultiRand()
    r_gpu.init_const(3, 0, 5)
    m = P.Pic1dpModules(gp)
    m.particle_init()
    m.field_init()
    m.particle_optimize_setup(rng=r_gpu, **sched)
    m.particle_set(0, st["x"], st["v"], st["p"], st["w"])
    m.interaction_collect_charge()
    m.field_solve_electric()
    t, np_gpu, e_gpu, flags = 0.0, [], [], 0
    for it in range(nsteps):
        for m.global_irk in (1, 2):
            m.interaction_push_particle()
            flags += m.particle_optimize(t)
            m.interaction_collect_charge()
            m.field_solve_electric()
        t += op.dt
        np_gpu.append(m.particle_get(0)["x"].size)
        e_gpu.append(m.gpu.field_energy())
    out = m.particle_get(0)
    f = m.gpu.get_field()
    m.particle_final()
    assert flags == 3 and np_gpu == np_trace
    assert rel_err(f["chargeden"], ref.rho) < 1e-10 and rel_err(f["electric"], ref.E) < 1e-10
    assert np.allclose(e_gpu, e_ref, rtol=1e-9)
    for k in ("x", "v", "p", "w"):
        assert rel_err(out[k], rs[k][:npar]) < 1e-11, k


def test_cpp_host_driver_with_optimisation_events_matches_oracle(tmp_path):
    """host/pic1dp_host with nmerge = nremove = nsplit = 1, all due in the same substep (t = 0.1): the product-side
    multirand (uniform + Gaussian streams), particle_load with unloaded tail markers (nparticle_init < nparticle_max,
    src/pic1dp_particle.F90:240-248) and the schedule of particle_optimize, against the oracle's replay."""
    import subprocess
    from pic1dp_b200 import build
    exe = build.build_host()
    cap, ninit, nx, nsteps = 300000, 200000, 192, 6
    out, mk = tmp_path / "energy.txt", tmp_path / "markers.bin"
    r = subprocess.run([exe, f"nparticle_max={cap}", f"nparticle_init={ninit}", f"nx={nx}", f"ntime_max={nsteps}",
                        "seed_type=1", "nmerge=1", "nremove=1", "nsplit=1", "opt_t0=0.0", "opt_dt=0.1", f"out={out}",
                        f"markers_out={mk}"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "marker optimisation" in r.stdout
    raw = np.fromfile(mk, dtype=np.float64)
    n_gpu = int(np.frombuffer(raw[:1].tobytes(), dtype=np.int64)[0])
    got = dict(zip(("x", "v", "p", "w"), raw[1:].reshape(4, n_gpu)))
    # ---- oracle replay ----
    op, _ = make_params(nx=nx)
    orc = O.Oracle(op)
    x, v, p, w = orc.particle_load(0, 3, 0, 5, cap, ninit)
    rng = O.MultiRand()
    rng.init_const(3, 0, 5)
    rng.real_array(cap)
    rng.real_array(cap)  # particle_load drew the whole local array twice (:180, :222)
    ref = OracleRun(op, [[dict(x=x, v=v, p=p, w=w)]])
    rs, npar = ref.st[0][0], ninit
    view = lambda m: [[{k: a[:m] for k, a in rs.items()}]]
    ref.st = view(npar)
    ref.init_field()
    energy, t = [ref.o.field_energy(ref.E)], 0.0
    done = False
    for it in range(nsteps):
        for irk in (1, 2):
            ref.st = view(npar)
            ref.push(irk)
            if irk == 2 and not done and t + op.dt >= 0.1:
                d = orc.dist_pertb_abs_v([rs["v"][:npar]], [rs["w"][:npar]], 128, VMAX)
                npar = orc.particle_merge(rs, npar, d, 0.1, VMAX)
                d = orc.dist_pertb_abs_v([rs["v"][:npar]], [rs["w"][:npar]], 128, VMAX)
                npar = orc.particle_remove(rs, npar, d, 0.1, 2, 0.9, rng, VMAX)
                d = orc.dist_pertb_abs_v([rs["v"][:npar]], [rs["w"][:npar]], 128, VMAX)
                npar = orc.particle_split(rs, npar, d, 0.1, 5, 0.1, rng, VMAX)
                done = True
            ref.st = view(npar)
            ref.collect_charge()
            ref.solve_field()
        t += op.dt
        energy.append(ref.o.field_energy(ref.E))
    assert done and n_gpu == npar and npar != ninit
    for k in ("x", "v", "p", "w"):
        assert rel_err(got[k], rs[k][:npar]) < 1e-11, k
    rows = np.loadtxt(out)
    assert np.allclose(rows[-1, 1], energy[-1], rtol=1e-9)
