"""Turns the ncu outputs a gpurun call brought back into the small text/JSON summaries kept under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.txt
  python tools/summarize_ncu.py kernel   gpurun_out/prof_conv_r1.ncu-rep > profiles/r1_conv_ncu.txt
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum.per_cycle_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def launches(path):
    lines = [l for l in open(path).read().splitlines() if l.startswith('"')]
    rows = list(csv.reader(io.StringIO("\n".join(lines))))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"# per-launch gpu__time_duration.sum (ns), {len(rows) - 1} launches; cold-cache & serialised: compare SHARES")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:28s} launches {c:4d}  total_ns {t:14.0f}  share {100 * t / tot:5.1f}%  mean_ns {t / c:10.0f}")


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                d[k] = f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip()
        out.append(d)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
