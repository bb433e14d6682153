#!/usr/bin/env python
"""Opcode evidence from the shipped binary: `cuobjdump -sass libcts_sm100a.so` -> per-kernel instruction counts of the
mnemonics that matter for this path (bulk TMA UBLKCP / mbarrier SYNCS / 128-bit LDG-STG / FP64 pipe / system-scope
atomics and fences of the peer-memory halo kernel).  No tensor-core opcodes are expected: nothing here is a contraction.

    python profiles/sass_opcodes.py > profiles/r2b_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "climatimesteppers.jl_b200", "libcts_sm100a.so")
GROUPS = [("UBLKCP", r"^UBLKCP"), ("UTMALDG/UTMASTG", r"^UTMA(LDG|STG)"), ("SYNCS(mbarrier)", r"^SYNCS"),
          ("LDG.128", r"^LDG\.E(\.\w+)*\.128"), ("STG.128", r"^STG\.E(\.\w+)*\.128"), ("LDG", r"^LDG"), ("STG", r"^STG"),
          ("LDS", r"^LDS"), ("STS", r"^STS"), ("SHFL", r"^SHFL"), ("BAR", r"^BAR"), ("DFMA/DADD/DMUL", r"^D(FMA|ADD|MUL)"),
          ("MUFU.RCP64H", r"^MUFU\.RCP64H"), ("ATOM/RED", r"^(ATOMG|ATOM|RED)"), ("*.SYS (system scope)", r"\.SYS"),
          ("MEMBAR", r"^MEMBAR"), ("DSMEM (MAPA / ST.*CLUSTER / UCGABAR)", r"^(MAPA|UCGABAR|ST\.E?\.?.*CLUSTER|STS.*CLUSTER)"), ("UTC*MMA/HMMA (tensor)", r"^(UTC\w*MMA|HMMA|HGMMA|LDTM|STTM)")]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kern, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kern[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][\w\.]*)", line)
        if m and cur:
            kern[cur]["_total"] += 1
            op = m.group(1)
            for name, pat in GROUPS:
                if re.search(pat, op):
                    kern[cur][name] += 1
    demangle = subprocess.run(["cu++filt"], input="\n".join(kern), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kern, demangle)) if len(demangle) == len(kern) else {k: k for k in kern}
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}   architectures: {arch}   kernels: {len(kern)}")
    tot = collections.Counter()
    for k in kern.values():
        tot.update(k)
    print("# whole library: " + ", ".join(f"{n} {tot[n]}" for n, _ in GROUPS))
    print()
    want = ("column_fused_stage", "tridiag_tile", "halo_peer", "column_final_update", "column_texp_march", "reduce_kernel",
            "reduce_ew_kernel", "axpy_n_vec", "ssprk_stage_vec", "advection_kernel", "column_residual")
    for k, c in kern.items():
        nm = names[k]
        i = nm.rfind(">(")
        short = nm[:i + 1] if i > 0 else nm.split("(")[0]
        short = re.sub(r"^void |\(anonymous namespace\)::|<unnamed>::|cts_\w+::|\((int|bool)\)", "", short)
        if not any(w in nm for w in want):
            continue
        print(f"{short[:110]:110s} inst {c['_total']:6d}  " + "  ".join(f"{n}={c[n]}" for n, _ in GROUPS if c[n]))


if __name__ == "__main__":
    main()
