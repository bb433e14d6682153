#!/usr/bin/env python
"""Per-instruction stall samples of an `ncu --page source --csv` dump (SASS view): totals per stall reason and the hottest
instructions of each kernel, with shared-memory bank-conflict (excessive wavefront) counts.
  python profiles/source_hotspots.py gpurun_out/r2_fused_source.csv.gz [top]"""
import csv
import gzip
import io
import sys


def main(path, top=25):
    raw = (gzip.open(path, "rt", errors="replace") if path.endswith(".gz") else open(path, errors="replace")).read()
    rows = list(csv.reader(io.StringIO(raw)))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    for k in kernels:
        h = k["hdr"]
        i_src, i_s, i_ex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        i_exc, i_wf = h.index("L1 Wavefronts Shared Excessive"), h.index("L1 Wavefronts Shared")
        stalls = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        tot = sum(int(r[i_s]) for r in k["rows"])
        inst = sum(int(r[i_ex]) for r in k["rows"])
        print("=" * 110)
        print(k["name"][:150])
        print(f"instructions (SASS lines) {len(k['rows'])}, warp instructions executed {inst}, stall samples {tot}, "
              f"shared wavefronts {sum(int(r[i_wf]) for r in k['rows'])} of which excessive (bank conflicts) {sum(int(r[i_exc]) for r in k['rows'])}")
        by = sorted(((sum(int(r[i]) for r in k["rows"]), h[i]) for i in stalls), reverse=True)
        print("stall reasons: " + ", ".join(f"{n} {100 * v / max(tot, 1):.1f}%" for v, n in by if v * 100 >= tot))
        print(f"{'samples':>8s} {'share':>6s} {'executed':>10s} {'excess_wf':>9s}  top stall        instruction")
        for r in sorted(k["rows"], key=lambda r: -int(r[i_s]))[:top]:
            best = max(stalls, key=lambda i: int(r[i]))
            print(f"{int(r[i_s]):8d} {100 * int(r[i_s]) / max(tot, 1):5.1f}% {int(r[i_ex]):10d} {int(r[i_exc]):9d}  {h[best][6:]:<16s} {r[i_src].strip()[:80]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
