#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rNN_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep      > profiles/rNN_ncu_full.txt

`launches`: per-kernel launch count, total / mean device time and SHARE of the captured launches (the
per-launch times of an ncu pass are cold-cache and serialised, so only shares are meaningful).
`full`: the metrics of the B200 profiling recipe for every captured launch of a `--set full` report.
"""
from __future__ import annotations

import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_bytes.sum", "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__cluster_dim_x",
]


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void\s+", "", name)
    return name if len(name) < 90 else name[:87] + "..."


def launches(path: str) -> None:
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    # torch.cuda._sleep: bench.py parks the GPU for ~10 ms so that the host queues its per-kernel timing loop ahead of it
    rows = [r for r in rows if "spin_kernel" not in r[4]]
    agg: "OrderedDict[str, list]" = OrderedDict()
    for r in rows:
        k = short(r[4])
        ns = float(r[-1])
        a = agg.setdefault(k, [0, 0.0, r[7], r[8]])
        a[0] += 1
        a[1] += ns
    total = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(rows)} launches, {total / 1e3:.1f} us of device time (cold-cache, serialised: read SHARES)")
    print(f"{'kernel':<92} {'n':>5} {'total us':>10} {'mean us':>9} {'share':>7}  block / grid")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:<92} {a[0]:>5} {a[1] / 1e3:>10.1f} {a[1] / a[0] / 1e3:>9.2f} {100 * a[1] / total:>6.1f}%  {a[2]} / {a[3]}")


def full(path: str) -> None:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: ncu --set full, {len(rows) - 2} launches")
    for r in rows[2:]:
        print(f"\n== {short(r[idx['Kernel Name']])}")
        for m in METRICS:
            if m in idx and r[idx[m]] != "":
                print(f"  {m:<66} {r[idx[m]]:>16} {units[idx[m]]}")
        try:
            rd, wr = float(r[idx["dram__bytes_read.sum"]]), float(r[idx["dram__bytes_write.sum"]])
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            tb = rd * scale[units[idx["dram__bytes_read.sum"]]] + wr * scale[units[idx["dram__bytes_write.sum"]]]
            print(f"  {'traffic (dram read + write)':<66} {tb / 1e6:>16.3f} MB")
        except Exception:
            pass


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
