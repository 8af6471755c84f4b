"""Groups an `ncu --metrics gpu__time_duration.sum --csv` launch list by (shortened) kernel name.
Usage: python tools/launch_summary.py launches.csv [top_n] [steps]"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("<unnamed>::", "").replace("at::native::", "").replace("native::", "")
    return name[:110]


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    steps = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    n = 0
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        a = agg[short(r[ki])]
        a[0] += 1
        a[1] += v
        total += v
        n += 1
    print("%d launches, %.3f ms total (%.3f ms / step over %g steps)" % (n, total, total / steps, steps))
    for name, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%8.3f ms %5.1f%% %6d x  %s" % (ms / steps, 100 * ms / total, c, name))


if __name__ == "__main__":
    main()
