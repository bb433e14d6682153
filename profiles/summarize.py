#!/usr/bin/env python
"""Turn the ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.
  python profiles/summarize.py launches gpurun_out/r1_launches.csv > profiles/r1_launches_summary.txt
  python profiles/summarize.py raw gpurun_out/r1_fused.ncu-rep   > profiles/r1_fused_stage_ncu.txt   (or the raw-page CSV)
  python profiles/summarize.py traffic gpurun_out/r2_fused_raw.csv stored_n1   -> profiles/fused_stage_traffic.json
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        if r is hdr or len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else (v * 1e3 if r[iu] in ("ms", "msecond") else v)  # -> us
        name = r[ik].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({path}); per-launch times are cold-cache and serialised:")
    print(f"# compare SHARES, not absolutes.  total {total / 1e3:.3f} ms over {sum(a[0] for a in agg.values())} launches")
    print(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {n:8d} {t:12.1f} {t / n:10.1f} {100 * t / total:6.1f}%")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct"]


def _raw_rows(path):
    """Rows of `ncu --page raw --csv`: from a .ncu-rep (converted here) or from a CSV already converted on the GPU box
    (a full capture with source is > 64 MiB, the size gpurun brings back)."""
    if path.endswith(".csv"):
        out = open(path, errors="replace").read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    i0 = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    return rows[i0:]


def traffic(path, key="stored_n1"):
    """profiles/fused_stage_traffic.json entry: mean dram read + write bytes per launch of the captured fused-stage launches
    (bench.py's roofline.traffic reads that file)."""
    import json
    import os
    rows = _raw_rows(path)
    hdr, units = rows[0], rows[1]
    sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    per = [float(r[ir].replace(",", "")) * sc[units[ir]] + float(r[iw].replace(",", "")) * sc[units[iw]] for r in rows[2:]]
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fused_stage_traffic.json")
    data = json.load(open(out)) if os.path.exists(out) else {}
    data[key] = {"traffic_bytes_per_launch": sum(per) / len(per), "per_launch": per, "launches": len(per), "source": os.path.basename(path),
                 "kernels": [r[hdr.index("Kernel Name")][:120] for r in rows[2:]]}
    json.dump(data, open(out, "w"), indent=1)
    print(json.dumps(data[key], indent=1))


def raw(path):
    rows = _raw_rows(path)
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none --import-source on  ({path})")
    for r in rows[2:]:
        print("-" * 100)
        print(r[hdr.index("Kernel Name")][:160])
        for i, h in enumerate(hdr):
            stall = h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")
            if h in WANT or stall:
                try:
                    if stall and float(r[i].replace(",", "")) < 0.2:
                        continue
                except ValueError:
                    pass
                print(f"  {h:90s} {units[i]:12s} {r[i]}")
        try:
            rd = float(r[hdr.index('dram__bytes_read.sum')].replace(",", ""))
            wr = float(r[hdr.index('dram__bytes_write.sum')].replace(",", ""))
            ur, uw = units[hdr.index('dram__bytes_read.sum')], units[hdr.index('dram__bytes_write.sum')]
            sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            print(f"  => traffic = dram read + write = {(rd * sc[ur] + wr * sc[uw]) / 1e9:.3f} GB per launch")
        except Exception:
            pass


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
