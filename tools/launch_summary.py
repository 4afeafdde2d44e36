"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (the SHARE of each kernel in
the profiled command; absolute times under ncu are cold-cache and serialised).

    python tools/launch_summary.py gpurun_out/launches.csv > profiles/rNN_launches_<what>.txt
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    col = {n: i for i, n in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")[:90]
        v, u = float(r[col["Metric Value"]]), r[col["Metric Unit"]]
        v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':92s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:92s} {v[0]:8d} {v[1]:10.1f} {v[1] / v[0]:8.2f} {100 * v[1] / tot:5.1f}%")
    print(f"{'TOTAL':92s} {sum(v[0] for v in agg.values()):8d} {tot:10.1f}")


if __name__ == "__main__":
    main()
