"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
    python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        try:
            v = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        unit = r[ix["Metric Unit"]]
        us = {"ns": v / 1e3, "us": v, "usecond": v, "ms": v * 1e3, "msecond": v * 1e3, "nsecond": v / 1e3}.get(unit, v)
        name = r[ix["Kernel Name"]]
        m = re.search(r"(\w+_kernel|\w+)<([^>]*)>\(", name)
        short = re.sub(r"^void\s+", "", re.sub(r"\(.*", "", name))
        short = short.replace("<unnamed>::", "").replace("ttasr::", "")[:80]
        key = (short, r[ix["Grid Size"]], r[ix["Block Size"]])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f"# launches: {sum(a[0] for a in agg.values())}, total device time {total / 1e3:.3f} ms "
          "(ncu per-launch times are cold-cache and serialised: compare shares)\n")
    print("| kernel | grid | block | launches | total ms | avg us | share |")
    print("|---|---|---|---:|---:|---:|---:|")
    for (short, grid, block), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{short}` | {grid} | {block} | {n} | {t / 1e3:.3f} | {t / n:.1f} | {100 * t / total:.2f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
