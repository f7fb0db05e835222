"""Reads `bench.py --dump-shapes` output and prints the step's tensor-core launches ranked by total isolated time.

    python tools/shape_table.py gpurun_out/shapes.jsonl [N]
"""
import json
import sys

MODES = {"0": "fwd", "1": "dgrad", "2": "wgrad"}


def main():
    rows = [json.loads(l) for l in open(sys.argv[1])]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    for r in rows:
        r["tot_us"] = r["us"] * r["count"]
    rows.sort(key=lambda r: -r["tot_us"])
    total = sum(r["tot_us"] for r in rows)
    print(f"{len(rows)} distinct shapes, {total / 1e3:.3f} ms per step")
    print(f"{'tot_us':>8} {'cnt':>4} {'us':>7} {'TF/s':>6}  tag / signature")
    for r in rows[:top]:
        s = r["sig"]
        if s[0] == "gemm":
            desc = (f"{MODES.get(s[1], s[1])} ext={s[2]} tile={s[3]} N={s[4]} M={s[5]} K/tap={s[6]} taps={s[7]} "
                    f"stride={s[8]} splits={s[9]} batch={s[10]} act={s[11]} odt={s[12]} res={s[13]} mask={s[14]}")
        else:
            desc = " ".join(s)
        tf = r["flops"] / (r["us"] * 1e-6) / 1e12 if r["us"] > 0 else 0.0
        print(f"{r['tot_us']:8.1f} {r['count']:4d} {r['us']:7.2f} {tf:6.1f}  {r['tag']}: {desc}")


if __name__ == "__main__":
    main()
