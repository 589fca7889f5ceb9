#!/usr/bin/env python
"""Condense a `summarize.py rep` markdown (one block of raw ncu metrics per captured launch) into one table row per launch:
duration, DRAM bytes read / written and the GB/s they amount to, DRAM %, tensor-pipe %, grid.
  python profiles/ncu_table.py profiles/r02d_prof_fwd.md [more.md ...]"""
import re
import sys

KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic")


def num(v, u):
    v = float(v.replace(",", ""))
    scale = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "ms": 1e3, "us": 1.0, "ns": 1e-3, "msecond": 1e3, "usecond": 1.0,
             "nsecond": 1e-3, "second": 1e6}
    return v * scale.get(u, 1.0)


def main(paths):
    print("| kernel | us | DRAM read MB | DRAM write MB | DRAM GB/s | DRAM % of peak | tensor pipe % | SM % | grid | regs | dyn smem KB |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for path in paths:
        for blk in open(path).read().split("### ")[1:]:
            name = blk.split("\n")[0].strip("`").replace("void ", "")
            vals = {}
            for line in blk.splitlines():
                m = re.match(r"\| (\S+) \| ([^|]+) \| ([^|]*) \|", line)
                if m and m.group(1) in KEYS:
                    vals[m.group(1)] = (m.group(2).strip(), m.group(3).strip())
            g = lambda k: num(*vals[k]) if k in vals else 0.0
            t, rd, wr = g(KEYS[0]), g(KEYS[1]), g(KEYS[2])
            print("| `%s` | %.1f | %.1f | %.1f | %.0f | %.1f | %.1f | %.1f | %d | %d | %.1f |" % (
                name, t, rd, wr, (rd + wr) / t * 1e3 if t else 0, g(KEYS[3]), g(KEYS[4]), g(KEYS[5]), g(KEYS[6]), g(KEYS[7]),
                g(KEYS[8]) * 1e3))


if __name__ == "__main__":
    main(sys.argv[1:])
