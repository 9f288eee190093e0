"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table (profiles/*.md).
usage: python tools/summarize_launches.py gpurun_out/launches.csv "title" > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r["Metric Unit"], 1.0)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("b200vc::<unnamed>::", "").replace("void ", "")
        rows.append((int(r["ID"]), name, r.get("Grid Size", ""), v))
    return rows


def main():
    path, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "launch list")
    rows = load(path)
    tot = collections.defaultdict(lambda: [0, 0.0])
    for _, n, _, v in rows:
        tot[n][0] += 1
        tot[n][1] += v
    T = sum(v[1] for v in tot.values())
    print(f"# {title}\n")
    print(f"{len(rows)} kernel launches, {T / 1e3:.1f} ms of GPU time (ncu serialises launches and runs them cold: compare SHARES, not absolutes).\n")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        if v[1] / T < 0.0005:
            continue
        print(f"| `{k[:70]}` | {v[0]} | {v[1] / 1e3:.2f} | {100 * v[1] / T:.1f}% | {v[1] / v[0]:.1f} |")


if __name__ == "__main__":
    main()
