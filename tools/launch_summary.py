"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (bench.py run under ncu).

    python tools/launch_summary.py profiles/r01_launches_bench.csv
"""
import collections
import csv
import sys


def main():
    lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            k = r["Kernel Name"].split("(")[0]
            agg[k][0] += 1
            agg[k][1] += float(r["Metric Value"]) / 1e6
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':58s} launches   total ms   share")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:58s} {v[0]:8d} {v[1]:10.2f} {100 * v[1] / tot:6.1f}%")


if __name__ == "__main__":
    main()
