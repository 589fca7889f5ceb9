#!/usr/bin/env python
"""Turn the artefacts that come back from the GPU box (gpurun_out/) into the tracked summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_H.csv  > profiles/rNN_launches_H.md
  python profiles/summarize.py rep gpurun_out/prof_fwd.ncu-rep     > profiles/rNN_prof_fwd.md

`launches`: per-kernel totals of an `ncu --metrics gpu__time_duration.sum` launch list (cold-cache, serialised: compare
SHARES, not absolutes).  `rep`: the roofline-relevant raw metrics of every launch in an `ncu --set full` capture.
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in r:
        if len(row) <= vi:
            continue
        v = float(row[vi].replace(",", ""))
        u = row[ui]
        v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
        name = re.sub(r"\(.*", "", row[ki]).replace("void ", "").replace("sivae::", "")
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    n = sum(a[0] for a in agg.values())
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.2f | %.1f%% |" % (k, c, ms, 100 * ms / tot))
    print("\ntotal %.1f ms over %d launches (ncu-serialised, cold cache)" % (tot, n))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg"]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
    ki = hdr.index("Kernel Name")
    for row in rows[2:]:
        print("### `%s`\n" % re.sub(r"\(.*", "", row[ki]))
        print("| metric | value | unit |\n|---|---:|---|")
        for w, i in idx:
            print("| %s | %s | %s |" % (w, row[i], units[i]))
        print()


if __name__ == "__main__":
    {"launches": launches, "rep": rep}[sys.argv[1]](sys.argv[2])
