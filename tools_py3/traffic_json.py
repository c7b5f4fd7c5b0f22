"""profiles/rNN_traffic.json from the ncu CSV of
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_push \
      -s 6 -c 4 --csv --log-file gpurun_out/traffic.csv python bench.py --steps 4 --warmup 3 --no-graph --no-cpu-baseline \
      --no-e2e --sustained-steps 0 --no-alt-arith
(steady-state launches of the two fused kernels at the bench size): measured DRAM bytes per marker per launch, which
bench.py scales by the markers of a launch for `roofline.traffic`.

  python tools_py3/traffic_json.py gpurun_out/traffic.csv MARKERS NX DEPOSIT_MODE > profiles/r02_traffic.json
"""
import csv
import json
import sys


def main():
    path, markers, nx, dep = sys.argv[1], int(float(sys.argv[2])), int(sys.argv[3]), int(sys.argv[4])
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    rows = [dict(zip(hdr, r)) for r in rows if r is not hdr and r[0].isdigit()]
    acc = {}
    for r in rows:
        name = r["Kernel Name"]
        irk = "irk2" if name.split("<")[1].split(",")[1].strip() == "1" else "irk1"   # k_push<DIST, IRK2, DEP, FUSED, CFG>
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "msecond": 1.0,
                 "usecond": 1e-3, "nsecond": 1e-6}.get(unit, 1.0)
        acc.setdefault(irk, {}).setdefault(r["Metric Name"], []).append(v * scale)
    out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                     "-k regex:k_push -s 6 -c 4 python bench.py --steps 4 --warmup 3 --no-graph ... at the BENCH size, round 2 "
                     "(tools_py3/traffic_json.py)",
           "markers": markers, "nx": nx, "deposit_mode": dep}
    for irk, alg in (("irk1", 56), ("irk2", 80)):
        m = acc[irk]
        rd = sum(m["dram__bytes_read.sum"]) / len(m["dram__bytes_read.sum"])
        wr = sum(m["dram__bytes_write.sum"]) / len(m["dram__bytes_write.sum"])
        out[irk] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_marker": (rd + wr) / markers,
                    "algorithmic_bytes_per_marker": alg, "ratio": (rd + wr) / markers / alg,
                    "ncu_duration_ms": sum(m["gpu__time_duration.sum"]) / len(m["gpu__time_duration.sum"])}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
