"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel family and per
(kernel, grid) shape.   python tools/launch_summary.py gpurun_out/launches.csv [--top 30]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 30
    lines = open(path).read().splitlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = list(csv.DictReader(lines[start:]))
    tot = 0.0
    fam = collections.defaultdict(lambda: [0, 0.0])
    shp = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        full = r["Kernel Name"]
        name = re.sub(r"\(.*", "", full).replace("void ", "")
        m = re.search(r"gemm_kernel<(\d+), (\d+)>", full)
        if m:
            name = f"gemm_kernel<BN={m.group(1)},mode={'FWD DGRAD WGRAD'.split()[int(m.group(2))]}>"
        t = float(r["Metric Value"]) / 1e3
        tot += t
        fam[name][0] += 1
        fam[name][1] += t
        key = f"{name} grid={r['Grid Size']}"
        shp[key][0] += 1
        shp[key][1] += t
    print(f"{len(rows)} launches, {tot / 1e3:.2f} ms of kernel time (cold-cache, serialised)")
    print("--- by kernel family")
    for k, (n, t) in sorted(fam.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t:9.1f} us {n:5d}x {100 * t / tot:5.1f}%  {k[:100]}")
    print("--- by (kernel, grid)")
    for k, (n, t) in sorted(shp.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{t:9.1f} us {n:5d}x {t / n:7.1f} us/launch {100 * t / tot:5.1f}%  {k[:110]}")


if __name__ == "__main__":
    main()
