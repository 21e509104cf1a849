#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python tools/ncu_launch_summary.py launches.csv [--top 40] > summary.md

Per-launch times under ncu are cold-cache and serialised: read the SHARES, not the absolutes."""
import csv
import io
import re
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"\(.*$", "", name)                   # drop the argument list
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<.*", lambda m: m.group(0) if len(m.group(0)) < 60 else m.group(0)[:57] + "...>", name)
    return name[:110]


def main():
    path = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = OrderedDict()
    total = 0.0
    n = 0
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(row["Metric Value"].replace(",", ""))
        if row.get("Metric Unit") == "us":
            ns *= 1e3
        k = short(row["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
        total += ns
        n += 1
    print(f"# {path}: {n} launches, {total / 1e6:.3f} ms of GPU time\n")
    print("| kernel | launches | total ms | avg us | share % |")
    print("|---|---|---|---|---|")
    for k, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"| `{k}` | {c} | {ns / 1e6:.3f} | {ns / c / 1e3:.1f} | {100 * ns / total:.1f} |")


if __name__ == "__main__":
    main()
