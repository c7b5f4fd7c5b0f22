"""world_size-2 gloo tests (CPU) of the N > 1 host-side logic: marker sharding by the PETSC_DECIDE block rule, the
unique-id broadcast plumbing bench.py uses, and that summing per-rank partial charge grids with an all-reduce
reproduces the single-rank density (the role of MPI_Allreduce at src/pic1dp_interaction.F90:132-133)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pic1dp_b200 as P
        from helpers import OracleRun, copy_state, make_params, synth_markers
        op, _ = make_params(nx=192)
        n = 50001
        st = synth_markers(op, n, seed=77)          # every rank generates the same global arrays
        lo, hi = P.petsc_decide(n, world, rank)     # ... and keeps its own contiguous block
        mine = {k: a[lo:hi].copy() for k, a in st.items()}
        # unique-id style broadcast of an opaque 128-byte token from rank 0
        tok = [bytes(range(128)) if rank == 0 else None]
        dist.broadcast_object_list(tok, src=0)
        assert tok[0] == bytes(range(128))
        # local deposit (charge2 of this rank), then all-reduce, then scale: :81-141
        run = OracleRun(op, [[mine]])
        c1, _ = run.o.deposit_species(mine["x"], mine["w"])
        c2 = torch.from_numpy(c1 * op.charge[0])
        dist.all_reduce(c2, op=dist.ReduceOp.SUM)
        rho = c2.numpy() * op.nx / op.lx
        counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([hi - lo]))
        if rank == 0:
            single = OracleRun(op, [[copy_state(st)]])
            single.collect_charge()
            err = float(np.max(np.abs(rho - single.rho)) / np.max(np.abs(single.rho)))
            q.put(("ok", err, int(sum(int(c) for c in counts))))
    except Exception as e:  # pragma: no cover
        q.put(("fail", repr(e), 0))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_allreduce_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    status, err, total = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
    assert status == "ok", err
    assert total == 50001
    assert err < 1e-12


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` is the CPU arm: must work without a GPU and print one JSON line."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-markers", "200000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "particle-steps/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
