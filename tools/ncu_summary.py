"""Summarises an `ncu --set full --csv --page raw` log into a markdown table + a {kernel: dram bytes/launch} JSON
(what bench.py's roofline.traffic reads).  Usage: python tools/ncu_summary.py raw.csv [out.md [traffic.json]]"""
import collections
import csv
import json
import re
import sys

COLS = [
    ("gpu__time_duration.sum", "ms", 1e-6),
    ("dram__bytes_read.sum", "dram rd MB", 1e-6),
    ("dram__bytes_write.sum", "dram wr MB", 1e-6),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %", 1),
    ("sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_elapsed", "smem pipe %", 1),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem LSU wavefronts %", 1),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %", 1),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1),
    ("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 2:]
    ki = hdr.index("Kernel Name")
    idx = [(hdr.index(c), label, scale) for c, label, scale in COLS if c in hdr]
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= ki:
            continue
        name = re.sub(r"^void |<unnamed>::|\(anonymous namespace\)::|\(.*", "", r[ki]).strip()
        # k_spconv_umma_dg<CIN, COUT, SPLIT, NSETS> / k_spconv_wgrad_umma<CIN, COUT, NG>: key by the layer shape only
        name = re.sub(r"^(k_spconv_umma_dg|k_spconv_wgrad_umma)<(\d+), (\d+)[^>]*>", r"\1<\2, \3>", name)
        vals = []
        for i, _, scale in idx:
            try:
                vals.append(float(r[i].replace(",", "")) * scale)
            except ValueError:
                vals.append(float("nan"))
        agg.setdefault(name, []).append(vals)
    lines = ["| kernel | launches | " + " | ".join(label for _, label, _ in idx) + " |",
             "|---|---|" + "---|" * len(idx)]
    traffic = {}
    for name, vs in agg.items():
        n = len(vs)
        mean = [sum(v[j] for v in vs) / n for j in range(len(idx))]
        lines.append("| `%s` | %d | " % (name, n) + " | ".join("%.3f" % m if j == 0 else "%.1f" % m
                                                                for j, m in enumerate(mean)) + " |")
        labels = [label for _, label, _ in idx]
        if "dram rd MB" in labels and "dram wr MB" in labels:
            traffic[name] = {"dram_bytes_per_launch": int((mean[labels.index("dram rd MB")] +
                                                           mean[labels.index("dram wr MB")]) * 1e6),
                             "ncu_ms_per_launch": round(mean[0], 4), "launches": n}
    out = "\n".join(lines)
    print(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "a").write(out + "\n")
    if len(sys.argv) > 3:
        json.dump(traffic, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
