"""Worker for tests/test_gpu_multi.py: one process per GPU (torchrun), particle decomposition by the PETSC_DECIDE
block rule, density all-reduce through the library's own NCCL communicator.  Rank 0 compares with the oracle's
emulated-rank run and prints MGPU_OK / MGPU_FAIL."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import pic1dp_b200 as P
    from helpers import OracleRun, make_params, rel_err, synth_markers

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    dep = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    p2p = len(sys.argv) > 2 and sys.argv[2] == "p2p"
    n, nx, nsteps = 400003, 256, 5
    op, gp = make_params(nx=nx, capacity=n, device=local, rank=rank, nranks=world, deposit_mode=dep)
    st = synth_markers(op, n, seed=99)
    lo, hi = P.petsc_decide(n, world, rank)
    g = P.Pic1dGpu(gp)
    uid = [g.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    g.comm_init(uid[0])
    if p2p:  # density all-reduce through peer memory instead of NCCL
        handles = [None] * world
        dist.all_gather_object(handles, g.p2p_export())
        g.p2p_import(handles)
    g.set_markers(0, *(np.ascontiguousarray(st[k][lo:hi]) for k in ("x", "v", "p", "w")))
    g.collect_charge()
    g.solve_field()
    g.step(nsteps)
    f = g.get_field()
    mk = g.get_markers(0)
    # every rank must hold the same replicated field
    E_all = [torch.zeros(nx, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(E_all, torch.from_numpy(f["electric"]))
    ok = all(torch.equal(E_all[0], e) for e in E_all)
    parts = [None] * world
    dist.all_gather_object(parts, {k: mk[k] for k in ("x", "v", "w")})
    c = g.counters()   # collectives of the time loop only (the optimisation phase below adds its own)
    # marker optimisation across ranks (src/pic1dp_particle.F90:356-522): the |w| histogram is all-reduced, the merge is
    # local to every rank; then the loop goes on with per-rank marker counts that differ
    dist_v = g.compute_dist_pertb_abs_v(128, 8.0)[0]
    pre = g.get_markers(0)
    np_after = g.particle_merge(0.3)[0]
    post = g.get_markers(0)
    g.collect_charge()
    g.solve_field()
    g.step(1)
    f2 = g.get_field()
    opt = [None] * world
    dist.all_gather_object(opt, dict(dist=dist_v, pre=pre, post=post, n=np_after))
    if rank == 0:
        states = []
        for r in range(world):
            a, b = P.petsc_decide(n, world, r)
            states.append({k: st[k][a:b].copy() for k in st})
        ref = OracleRun(op, [states])
        ref.init_field()
        for _ in range(nsteps):
            ref.step()
        e_rho, e_E = rel_err(f["chargeden"], ref.rho), rel_err(f["electric"], ref.E)
        e_mk = max(rel_err(parts[r][k], ref.st[0][r][k]) for r in range(world) for k in ("x", "v", "w"))
        ncoll = 2 * nsteps + 1
        coll_ok = (c.p2p_allreduces == ncoll and c.p2p_timeouts == 0 and c.nccl_calls == 0) if p2p else c.nccl_calls == ncoll
        # optimisation phase against the oracle's emulated ranks
        d_ref = ref.o.dist_pertb_abs_v([ref.st[0][r]["v"] for r in range(world)], [ref.st[0][r]["w"] for r in range(world)],
                                       128, 8.0)
        e_dist = rel_err(opt[0]["dist"], d_ref)
        dist_same = all(np.array_equal(opt[0]["dist"], o["dist"]) for o in opt)
        merge_ok = True
        for r in range(world):   # the merge itself is bit-exact given the rank's own pre-state and the all-reduced dist
            m = {k: opt[r]["pre"][k].copy() for k in ("x", "v", "p", "w")}
            n_ref = ref.o.particle_merge(m, m["x"].size, opt[r]["dist"], 0.3, 8.0)
            merge_ok &= n_ref == opt[r]["n"] and n_ref < m["x"].size
            merge_ok &= all(np.array_equal(m[k][:n_ref], opt[r]["post"][k]) for k in ("x", "v", "p", "w"))
            ref.st[0][r] = {k: m[k][:n_ref].copy() for k in ("x", "v", "p", "w")}
            for k in ("xb", "vb", "wb"):
                ref.st[0][r][k] = np.zeros(n_ref)
        ref.collect_charge()
        ref.solve_field()
        ref.step()
        e_E2 = rel_err(f2["electric"], ref.E)
        good = ok and e_rho < 1e-12 and e_E < 1e-12 and e_mk < 1e-12 and coll_ok
        good = good and e_dist < 1e-12 and dist_same and merge_ok and e_E2 < 1e-10
        print(("MGPU_OK" if good else "MGPU_FAIL"), f"world={world} rho={e_rho:.2e} E={e_E:.2e} markers={e_mk:.2e} "
              f"replicated_equal={ok} nccl_calls={c.nccl_calls} p2p={c.p2p_allreduces} timeouts={c.p2p_timeouts} "
              f"dist={e_dist:.2e} dist_same={dist_same} merge_ok={merge_ok} E_after_merge={e_E2:.2e}", flush=True)
    g.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
