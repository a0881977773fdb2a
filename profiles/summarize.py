#!/usr/bin/env python
"""Turn the ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv profiles/r1_launches.md
    python profiles/summarize.py raw gpurun_out/raw_gemm.csv [more.csv ...] profiles/r1_kernels.md

`launches`: the `--metrics gpu__time_duration.sum` launch list of one bench.py run -> per-kernel launch count, total
device time and share (cold-cache, serialised: only the SHARES are meaningful).
`raw`: `ncu -i X.ncu-rep --page raw --csv` of a `--set full` capture -> one row per captured launch with duration, DRAM
bytes, DRAM %, tensor-pipe %, L2 hit rate, registers, clock.
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("void ", "").replace("dfd::", "")


def launches(src, dst, skip_pack=True, last_predicts=0):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    if last_predicts:
        # a predict starts with the patch extraction kernel: keep the last N whole predicts of the run
        starts = [i for i, r in enumerate(rows) if i > 0 and "patchify_kernel" in r[ki]]
        rows = [hdr] + rows[starts[-last_predicts]:]
    agg = OrderedDict()
    for r in rows[1:]:
        k = short(r[ki])
        a = agg.setdefault(k, [0, 0.0, r[gi]])
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) / 1e3
    total = sum(v[1] for v in agg.values())
    with open(dst, "w") as fh:
        fh.write("| kernel | launches | total us | share | grid (first launch) |\n|---|---:|---:|---:|---|\n")
        for k, (n, us, grid) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write("| `%s` | %d | %.1f | %.1f%% | %s |\n" % (k, n, us, 100 * us / total, grid))
        fh.write("| **total** | %d | %.1f | 100%% | |\n" % (sum(v[0] for v in agg.values()), total))


COLS = [
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "rd MB"),
    ("dram__bytes_write.sum", "wr MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("sm__cycles_elapsed.avg.per_second", "GHz"),
]


def to_mb(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)


def raw(srcs, dst):
    with open(dst, "w") as fh:
        fh.write("| kernel | " + " | ".join(c[1] for c in COLS) + " |\n|---|" + "---:|" * len(COLS) + "\n")
        for src in srcs:
            rows = list(csv.reader(open(src)))
            hdr, units = rows[0], rows[1]
            ki = hdr.index("Kernel Name")
            for r in rows[2:]:
                vals = []
                for name, label in COLS:
                    if name not in hdr:
                        vals.append("-")
                        continue
                    i = hdr.index(name)
                    v = r[i]
                    if label.endswith("MB"):
                        vals.append("%.1f" % to_mb(v, units[i]))
                    elif label == "us":
                        f = float(v.replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[i], 1.0)
                        vals.append("%.1f" % f)
                    else:
                        try:
                            vals.append("%.4g" % float(v.replace(",", "")))
                        except ValueError:
                            vals.append(v)
                k = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("dfd::", "")
                fh.write("| `%s` | %s |\n" % (k, " | ".join(vals)))


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif mode == "launches_last":  # launches_last N src.csv dst.md: only the last N predicts of the run
        launches(sys.argv[3], sys.argv[4], last_predicts=int(sys.argv[2]))
    else:
        raw(sys.argv[2:-1], sys.argv[-1])
