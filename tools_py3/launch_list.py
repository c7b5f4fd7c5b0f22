"""Turn the CSV of `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv <cmd>` into the
launch-list summary kept under profiles/ (per-kernel totals and shares, first launches in order).

  python tools_py3/launch_list.py gpurun_out/launches.csv "<command that was profiled>" [plain_run.json] > profiles/rNN_launch_list.md
"""
import collections
import csv
import json
import sys


def main():
    path, cmd = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ki, vi, ii, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("ID"), h.index("Metric Unit")
    bi, gi = h.index("Block Size"), h.index("Grid Size")
    launches = []
    for r in rows[1:]:
        t = float(r[vi].replace(",", ""))
        unit = r[ui]
        ms = t / 1e6 if unit in ("ns", "nsecond") else t / 1e3 if unit in ("us", "usecond") else t
        launches.append((int(r[ii]), r[ki], ms, r[bi], r[gi]))
    tot = collections.OrderedDict()
    for _, k, ms, _, _ in launches:
        a = tot.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(v[1] for v in tot.values())
    print("# ncu launch list\n")
    print(f"Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv {cmd}`, run directly after the same "
          "command exited 0 without ncu. Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.\n")
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {ms:.3f} | {100 * ms / total:.1f}% |")
    print("\nFirst 16 launches in order (name, ms, block, grid):\n")
    for _, k, ms, b, g in launches[:16]:
        print(f"- `{k}` {ms:.4f} ms {b} {g}")
    push = sum(ms for k, (n, ms) in tot.items() if "k_push" in k)
    print(f"\nShare check: the fused particle kernels are {100 * push / total:.1f}% of device time under ncu.", end=" ")
    if len(sys.argv) > 3:
        d = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
        i1, i2 = d["roofline_detail"]["irk1"]["ms_per_launch"], d["roofline"]["ms_per_launch"]
        print(f"In the plain run of the same command (CUDA events inside the timed region) they are "
              f"({i2:.3f}+{i1:.3f})/{d['ms_per_step']:.3f} = {100 * (i1 + i2) / d['ms_per_step']:.1f}% of the step "
              f"(value {d['value']:.3e} particle-steps/s, step {100 * d['roofline_detail']['step']['frac']:.1f}% of the measured HBM "
              f"roofline, irk=2 kernel {100 * d['roofline']['frac']:.1f}%).", end=" ")
    # steady state: launches after the one-time loader / initial deposit
    steady = [l for l in launches if l[0] >= next((x[0] for x in launches if "k_push" in x[1]), 0)]
    sp = sum(ms for _, k, ms, _, _ in steady if "k_push" in k)
    st = sum(ms for _, k, ms, _, _ in steady if not any(t in k for t in ("k_diag", "k_load_markers")))
    print(f"\n\nSteady state (launches from the first fused kernel on, diagnostics excluded): fused particle kernels "
          f"{sp:.3f} ms of {st:.3f} ms = {100 * sp / st:.1f}% under ncu.")


if __name__ == "__main__":
    main()
