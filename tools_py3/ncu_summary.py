"""Summarise an .ncu-rep (read with `ncu -i`): per-kernel headline metrics + SASS opcode mix + stall reasons.

  python tools_py3/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx.md
"""
import collections
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
       "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
       "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    print(f"# ncu summary of {rep}\n")
    for r in raw[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"## launch {d.get('ID')}: {d.get('Kernel Name')}\n")
        for k in RAW:
            if k in d:
                print(f"- {k} = {d[k]} {u[k]}")
        print()
    src = run([rep, "--page", "source", "--csv"])
    blocks = src.split('"Kernel Name",')[1:]
    for b in blocks:
        lines = b.split("\n")
        name = lines[0].strip().strip('",')
        rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
        if not rows:
            continue
        h = rows[0]
        ix = {n: i for i, n in enumerate(h)}
        ops = collections.Counter()
        samples = collections.Counter()
        stalls = collections.Counter()
        total_inst = 0
        stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        top = []
        for r in rows[1:]:
            if len(r) < len(h):
                continue
            sass = r[ix["Source"]].strip()
            op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
            op = ".".join(op.split(".")[:2])
            n = int(r[ix["Instructions Executed"]] or 0)
            s = int(r[ix["# Samples"]] or 0)
            ops[op] += n
            samples[op] += s
            total_inst += n
            for c in stall_cols:
                stalls[c] += int(r[ix[c]] or 0)
            top.append((s, sass))
        print(f"## SASS profile: {name}\n")
        print(f"warp-instructions executed: {total_inst}\n")
        print("| opcode | warp-inst | share | stall samples |\n|---|---|---|---|")
        for op, n in ops.most_common(28):
            print(f"| {op} | {n} | {100.0 * n / max(total_inst, 1):.1f}% | {samples[op]} |")
        tot = sum(stalls.values())
        print("\n| stall reason | samples | share |\n|---|---|---|")
        for c, n in stalls.most_common(10):
            print(f"| {c} | {n} | {100.0 * n / max(tot, 1):.1f}% |")
        print("\nTop sampled instructions:\n")
        for s, sass in sorted(top, reverse=True)[:12]:
            print(f"- {s}: `{sass}`")
        print()


if __name__ == "__main__":
    main()
