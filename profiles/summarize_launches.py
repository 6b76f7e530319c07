"""Launch list (ncu --metrics gpu__time_duration.sum --csv) -> markdown: per-kernel totals and shares.
Usage: python profiles/summarize_launches.py <launches.csv> <steps> [title]"""
import csv
import io
import sys
from collections import OrderedDict


def main():
    path, steps = sys.argv[1], int(sys.argv[2])
    title = sys.argv[3] if len(sys.argv) > 3 else path
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    tot = OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        t = float(r["Metric Value"]) / (1e3 if r["Metric Unit"] == "ns" else 1.0)
        n, s = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, s + t)
    total = sum(s for _, s in tot.values())
    print(f"# {title}\n")
    print(f"total {total:.1f} us over {steps} steps = {total / steps:.1f} us per step (serialised, cold cache)\n")
    for name, (n, s) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"* `{name[:70]}`: {n} launches, {s:.1f} us, {100 * s / total:.1f} %")


if __name__ == "__main__":
    main()
