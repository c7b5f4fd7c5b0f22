"""Oracle comparisons at the sizes BASELINE.json's configs state (the round-1 suite compared at <= 4e6 markers):

  configs[1]  bump-on-tail, 1e7 markers, nx = 256: cell index bit-exact at every substep of 10 steps; rho, E within
              1e-12 of max; x, v, w within 1e-12 after the run.
  configs[2]  Landau-damping Maxwellian, nx = 4096 (deposition-scatter stress), 1e7 markers, 3 steps.
  configs[3]  one RK substep at the bench state (1e8 markers, nx = 1024) with a prescribed E: x, v bit-exact,
              w <= 1e-14, rho <= 1e-12.
The oracle runs as emulated MPI ranks on a thread pool (tests/helpers.ChunkedOracleRun)."""
import os

import numpy as np
import pytest

import pic1dp_b200 as P
from helpers import ChunkedOracleRun, copy_state, make_params, rel_err, synth_markers

pytestmark = pytest.mark.gpu

NCHUNK = max(4, min(32, os.cpu_count() or 4))
TOL_SUM = 1e-12
TOL_W = 1e-14


@pytest.mark.parametrize("dep", [P.DEPOSIT_AUTO, P.DEPOSIT_FIXED])
def test_config1_bump_on_tail_1e7_markers_nx256(dep):
    n = 10_000_000
    op, gp = make_params(nx=256, capacity=n, deposit_mode=dep)
    st = synth_markers(op, n, seed=101)
    ref = ChunkedOracleRun(op, copy_state(st), NCHUNK)
    with P.Pic1dGpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        ref.init_field()
        g.collect_charge()
        g.solve_field()
        f = g.get_field()
        assert rel_err(f["chargeden"], ref.rho) < TOL_SUM and rel_err(f["electric"], ref.E) < TOL_SUM
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
                gix = g.get_shape_x(0)[0]
                assert np.array_equal(gix, ref.cell_index()), (it, irk)   # bit-exact cell index, all 1e7 markers
        out = g.get_markers(0)
    for k in ("x", "v", "w"):
        assert rel_err(out[k], ref.full[k]) < 1e-12, k
    assert np.array_equal(out["p"], ref.full["p"])


@pytest.mark.parametrize("dep", [P.DEPOSIT_AUTO, P.DEPOSIT_FIXED])
def test_config2_landau_nx4096_1e7_markers(dep):
    n = 10_000_000
    kw = dict(nx=4096, capacity=n, deposit_mode=dep, iptcldist=0, density=[1.0], v0=[0.0], lx=4.0 * np.pi)
    op, gp = make_params(**kw)
    st = synth_markers(op, n, seed=102)
    ref = ChunkedOracleRun(op, copy_state(st), NCHUNK)
    with P.Pic1dGpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        ref.init_field()
        g.collect_charge()
        g.solve_field()
        for it in range(3):
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
                assert np.array_equal(g.get_shape_x(0)[0], ref.cell_index()), (it, irk)
        out = g.get_markers(0)
    for k in ("x", "v", "w"):
        assert rel_err(out[k], ref.full[k]) < 1e-12, k


def test_config3_one_substep_at_the_bench_state_1e8_markers():
    """The state bench.py times (1e8 markers, nx = 1024, bump-on-tail delta-f), both RK substeps with a prescribed E:
    the fused kernels against the oracle's push + wrap + deposit."""
    n = 100_000_000
    op, gp = make_params(nx=1024, capacity=n)
    st = synth_markers(op, n, seed=103)
    E = 1e-3 * np.sin(2 * np.pi * np.arange(op.nx) / op.nx + 0.3)
    ref = ChunkedOracleRun(op, st, NCHUNK)      # the oracle advances `st` in place; the GPU gets its copy first
    with P.Pic1dGpu(gp) as g:
        g.set_markers(0, st["x"], st["v"], st["p"], st["w"])
        for irk in (1, 2):
            g.set_field(electric=E)
            ref.E = E.copy()
            ref.push(irk)
            ref.collect_charge()          # wraps x, deposits
            g.push(irk)                   # fused: push + wrap + deposit
            g.collect_charge()
            out = g.get_markers(0, want=("x", "v", "w"))
            assert np.array_equal(out["x"], ref.full["x"]), irk
            assert np.array_equal(out["v"], ref.full["v"]), irk
            assert rel_err(out["w"], ref.full["w"]) < TOL_W, irk
            assert rel_err(g.get_field()["chargeden"], ref.rho) < TOL_SUM, irk
            ref.full["w"][:] = out["w"]   # continue from identical state so irk = 2 isolates one substep
            del out
