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
        c = g.counters()
        ncoll = 2 * nsteps + 1
        coll_ok = (c.p2p_allreduces == ncoll and c.p2p_timeouts == 0 and c.nccl_calls == 0) if p2p else c.nccl_calls == ncoll
        good = ok and e_rho < 1e-12 and e_E < 1e-12 and e_mk < 1e-12 and coll_ok
        print(("MGPU_OK" if good else "MGPU_FAIL"), f"world={world} rho={e_rho:.2e} E={e_E:.2e} markers={e_mk:.2e} "
              f"replicated_equal={ok} nccl_calls={c.nccl_calls} p2p={c.p2p_allreduces} timeouts={c.p2p_timeouts}", flush=True)
    g.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
