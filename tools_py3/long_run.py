"""Long-run acceptance on N GPUs: bump-on-tail at BASELINE.json configs[4] scale (1e9 markers total, nx = 8192 at 8
GPUs), energy recorded every 10 steps through pic1dp_gpu_output_field, growth rate fitted like tools/runinfo.py -gr.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools_py3/long_run.py \
      --markers-per-gpu 1.25e8 --nx 8192 --tmax 60 --out profiles/r01_long_run_c5.json
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--markers-per-gpu", type=float, default=1.25e8)
    ap.add_argument("--nx", type=int, default=8192)
    ap.add_argument("--tmax", type=float, default=60.0)
    ap.add_argument("--fit", type=float, nargs=2, default=[20.0, 50.0])
    ap.add_argument("--deposit", type=int, default=0)
    ap.add_argument("--arith", default="strict", choices=["strict", "tolerance"])
    ap.add_argument("--rng", default="host", choices=["host", "kiss64"],
                    help="host: numpy markers uploaded with set_markers; kiss64: particle_load with the device KISS64 stream")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import pic1dp_b200 as P
    from bench import fill_markers
    from tools_py3.runinfo import growthrate_energy_fit

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    n = int(args.markers_per_gpu)
    gp = P.default_params(nx=args.nx, capacity=n, device=local, rank=rank, nranks=world, deposit_mode=args.deposit,
                          arith_mode=1 if args.arith == "tolerance" else 0)
    g = P.Pic1dGpu(gp)
    if world > 1:
        uid = [g.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        g.comm_init(uid[0])
        handles = [None] * world
        dist.all_gather_object(handles, g.p2p_export())
        g.p2p_import(handles)
    if args.rng == "kiss64":
        from bench import rank_seeds
        g.load_markers_kiss64(0, n, rank_seeds(rank), 0, n, n * world, v_max=8.0)
    else:
        x, v, p, w = (np.empty(n) for _ in range(4))
        fill_markers(x, v, p, w, gp.lx, seed=4321 + rank)
        p *= 1.0 / world   # fill_markers normalises p for n markers; the plasma holds n * world of them
        w *= 1.0 / world
        g.set_markers(0, x, v, p, w)
        del x, v, p, w
    g.collect_charge()
    g.solve_field()
    nout = int(round(args.tmax / (10 * gp.dt)))
    t, en = [0.0], [g.output_field()[0]]
    g.sync()
    t0 = time.perf_counter()
    for k in range(1, nout + 1):
        g.step(10)
        en.append(g.output_field()[0])
        t.append(10 * gp.dt * k)
    g.sync()
    wall = time.perf_counter() - t0
    c = g.counters()
    if rank == 0:
        gamma = growthrate_energy_fit(np.array(t), np.array(en), args.fit[0], args.fit[1]) / 2.0
        res = {"markers_total": n * world, "n_gpus": world, "nx": args.nx, "steps": nout * 10, "tmax": args.tmax,
               "fit_window": args.fit, "gamma_fit": gamma, "gamma_analytic": 0.0838311,
               "rel_dev": gamma / 0.0838311 - 1.0, "wall_s_incl_outputs": wall,
               "particle_steps_per_s_wall": n * world * nout * 10 / wall, "oob_markers": int(c.oob_markers),
               "p2p_allreduces": int(c.p2p_allreduces), "p2p_timeouts": int(c.p2p_timeouts), "nccl_calls": int(c.nccl_calls),
               "deposit_mode": int(c.deposit_mode), "arith_mode": args.arith, "rng": args.rng,
               "graph_replays": int(c.graph_replays), "t": t, "energy": en}
        print(json.dumps({k: v for k, v in res.items() if k not in ("t", "energy")}))
        if args.out:
            json.dump(res, open(args.out, "w"))
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
