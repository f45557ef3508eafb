"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the step).
    python profiles/ncu_summary.py gpurun_out/launches.csv [steps] > profiles/launches_rNN.txt
"""
import collections
import csv
import re
import sys


def main(path, steps=1):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
        k = re.sub(r"\(.*", "", row["Kernel Name"])
        k = re.sub(r"^void ", "", k)[:90]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if "<unnamed>::" in k or "b2a" in k)
    print("# %s: %d launches over %d step(s); total %.1f us/step; libb2a kernels %.1f us/step (%.1f %%)" %
          (path, n, steps, tot / steps, ours / steps, 100 * ours / max(tot, 1e-9)))
    print("# per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes")
    print("%12s %8s %7s  kernel" % ("us/step", "launches", "share"))
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%12.1f %8.1f %6.1f%%  %s" % (t / steps, c / steps, 100 * t / tot, k))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
