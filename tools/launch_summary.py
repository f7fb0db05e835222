"""Summarises an `ncu --csv` launch list (one row per launch and metric): device time per kernel family and per
(kernel, grid) shape; with `dram__bytes_read.sum` / `dram__bytes_write.sum` in the capture also the DRAM traffic per
family, and `--json out.json` writes the per-family totals (bench.py reads profiles/r02_dram_traffic.json for
`roofline.traffic`).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py
    python tools/launch_summary.py gpurun_out/launches.csv [--top 30] [--json profiles/r02_dram_traffic.json]
"""
import collections
import csv
import json
import re
import sys

_UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3,
         "second": 1e6, "ns": 1e-3, "us": 1.0, "ms": 1e3}


def family(full: str) -> str:
    name = re.sub(r"\(.*", "", full).replace("void ", "")
    m = re.search(r"gemm_kernel<(\d+), (\d+)", full)
    if m:
        name = f"toist::gemm_kernel<BN={m.group(1)},{'FWD DGRAD WGRAD'.split()[int(m.group(2))]}>"
    return name


def main():
    path = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
    out_json = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    launches = collections.OrderedDict()  # ID -> {name, grid, time_us, rd, wr}
    for r in csv.DictReader(lines[start:]):
        d = launches.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"], "time_us": 0.0, "rd": 0.0, "wr": 0.0})
        val = float(r["Metric Value"].replace(",", "")) * _UNIT.get(r.get("Metric Unit", ""), 1.0)
        metric = r["Metric Name"]
        if metric.startswith("gpu__time_duration"):
            d["time_us"] = val
        elif metric.startswith("dram__bytes_read"):
            d["rd"] = val
        elif metric.startswith("dram__bytes_write"):
            d["wr"] = val
    rows = list(launches.values())
    tot = sum(d["time_us"] for d in rows)
    have_dram = any(d["rd"] or d["wr"] for d in rows)
    fam = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    shp = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for d in rows:
        f = family(d["name"])
        for table, key in ((fam, f), (shp, f"{f} grid={d['grid']}")):
            t = table[key]
            t[0] += 1
            t[1] += d["time_us"]
            t[2] += d["rd"]
            t[3] += d["wr"]
    print(f"{len(rows)} launches, {tot / 1e3:.2f} ms of kernel time (cold-cache, serialised)"
          + (f", DRAM read {sum(d['rd'] for d in rows) / 1e9:.2f} GB + write {sum(d['wr'] for d in rows) / 1e9:.2f} GB" if have_dram else ""))
    print("--- by kernel family")
    for k, (n, t, rd, wr) in sorted(fam.items(), key=lambda kv: -kv[1][1])[:top]:
        extra = f"  dram {rd / 1e6:9.1f} MB rd {wr / 1e6:9.1f} MB wr" if have_dram else ""
        print(f"{t:9.1f} us {n:5d}x {100 * t / tot:5.1f}%{extra}  {k[:100]}")
    print("--- by (kernel, grid)")
    for k, (n, t, rd, wr) in sorted(shp.items(), key=lambda kv: -kv[1][1])[:top]:
        extra = f"  dram/launch {rd / n / 1e6:7.2f}+{wr / n / 1e6:7.2f} MB" if have_dram else ""
        print(f"{t:9.1f} us {n:5d}x {t / n:7.1f} us/launch {100 * t / tot:5.1f}%{extra}  {k[:110]}")
    if out_json:
        g = [v for k, v in fam.items() if "gemm_kernel" in k]
        a = [v for k, v in fam.items() if "attn_" in k]
        res = {"source": path, "launches": len(rows), "kernel_time_us": tot,
               "gemm_family": {"launches": sum(v[0] for v in g), "time_us": sum(v[1] for v in g),
                               "dram_read_bytes": sum(v[2] for v in g), "dram_write_bytes": sum(v[3] for v in g)},
               "attention_family": {"launches": sum(v[0] for v in a), "time_us": sum(v[1] for v in a),
                                    "dram_read_bytes": sum(v[2] for v in a), "dram_write_bytes": sum(v[3] for v in a)},
               "all": {"dram_read_bytes": sum(d["rd"] for d in rows), "dram_write_bytes": sum(d["wr"] for d in rows)}}
        with open(out_json, "w") as f:
            json.dump(res, f, indent=1)
        print("wrote", out_json)


if __name__ == "__main__":
    main()
