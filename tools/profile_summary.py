#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/profile_summary.py launches gpurun_out/launches_r1b.csv > profiles/r1_launches_summary.md
    python tools/profile_summary.py full gpurun_out/prof_conv_r1b.ncu-rep > profiles/r1_conv_tc_ncu_full.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe insts"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM read"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    for r in rows:
        k = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            v *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            v *= 1e6
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.3f | %.1f%% |" % (k, n, t / 1e6, 100 * t / total))
    print("\n%d launches, %.2f ms of kernel time (ncu: cold cache, serialised; compare SHARES, not absolutes)" % (len(rows), total / 1e6))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    have = [(k, n) for k, n in KEYS if k in hdr]
    print("| # | kernel | " + " | ".join(n for _, n in have) + " |")
    print("|---|---|" + "---|" * len(have))
    for i, d in enumerate(data):
        name = d[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
        cells = []
        for k, _ in have:
            j = hdr.index(k)
            v = d[j]
            try:
                v = "%.4g" % float(v.replace(",", ""))
            except ValueError:
                pass
            cells.append("%s %s" % (v, units[j]))
        print("| %d | `%s` | %s |" % (i, name, " | ".join(cells)))


def traffic(path, planes, batch):
    """DRAM bytes (read + write) of every conv_tc 3x3 launch in an `ncu --set full` capture of ONE step -> JSON."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = []
    for d in data:
        if "conv_tc_kernel" not in d[hdr.index("Kernel Name")]:
            continue
        per.append({"read": float(d[ir]) * scale[units[ir]], "write": float(d[iw]) * scale[units[iw]],
                    "ms": float(d[it]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "msecond": 1.0, "usecond": 1e-3}[units[it]]})
    print(json.dumps({"planes": int(planes), "batch": int(batch), "launches": len(per),
                      "dram_bytes_per_step": sum(q["read"] + q["write"] for q in per), "per_launch": per}))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
