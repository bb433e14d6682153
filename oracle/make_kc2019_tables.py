#!/usr/bin/env python
"""Generate oracle/kc2019_tables.json: the rational coefficient tables of ARK437L2SA1 and ARK548L2SA2 (Kennedy &
Carpenter 2019, Table 8) as the reference states them (src/solvers/imex_tableaus.jl:883-1066).

~200 fifteen-digit rationals are not re-typed by hand: this script reads the reference source where it lies
(/root/reference, only available in the build container) and records every `name[i, j] = num // den` / `name[i] = num // den`
assignment of the two constructors, then applies the constructors' structural rules (first column = second column,
b_imp = b_exp, stiffly-accurate last row, ...) exactly as written there.  The committed JSON is test infrastructure of
the oracle (oracle/cts_ref.py `tableau`); the product's own table lives in climatimesteppers.jl_b200/tableaus.py and was
transcribed independently -- tests/test_tableaus_cpu.py compares the two.

    python oracle/make_kc2019_tables.py [/root/reference]
"""
import json
import os
import re
import sys
from fractions import Fraction

HERE = os.path.dirname(os.path.abspath(__file__))


def parse(src, name, s):
    a = src.index(f"function IMEXTableau(::{name})")
    body = src[a:src.index("IMEXTableau(;", a + 30)]
    tabs = {"a_exp": {}, "a_imp": {}, "b": {}, "c": {}}
    gamma = Fraction(*map(int, re.search(r"γ = (\d+) // (\d+)", body).groups()))
    for m in re.finditer(r"^\s*(a_exp|a_imp)\[(\d+), (\d+)\] = (-?\d+)(?: // (\d+))?\s*$", body, re.M):
        tabs[m.group(1)][(int(m.group(2)), int(m.group(3)))] = Fraction(int(m.group(4)), int(m.group(5) or 1))
    for m in re.finditer(r"^\s*(b|c)\[(\d+)\] = (-?\d+)(?: // (\d+))?\s*$", body, re.M):
        tabs[m.group(1)][int(m.group(2))] = Fraction(int(m.group(3)), int(m.group(4) or 1))
    A_exp = [[Fraction(0)] * s for _ in range(s)]
    A_imp = [[Fraction(0)] * s for _ in range(s)]
    b = [Fraction(0)] * s
    c = [Fraction(0)] * s
    for i in range(2, s + 1):
        A_imp[i - 1][i - 1] = gamma
    for (i, j), v in tabs["a_imp"].items():
        A_imp[i - 1][j - 1] = v
    for (i, j), v in tabs["a_exp"].items():
        A_exp[i - 1][j - 1] = v
    for i, v in tabs["b"].items():
        b[i - 1] = v
    for i, v in tabs["c"].items():
        c[i - 1] = v
    # structural rules, in the order the constructor applies them
    if name == "ARK548L2SA2":
        A_exp[7][1] = A_exp[7][0]                      # a_exp[8, 2] = a_exp[8, 1]
    for i in range(2, s + 1):
        A_imp[i - 1][0] = A_imp[i - 1][1]              # a_imp[i, 1] = a_imp[i, 2]
    b[0] = b[1]
    if name == "ARK548L2SA2":
        b[s - 1] = gamma                               # b[8] = γ
        for i in range(s):
            A_imp[s - 1][i] = b[i]                     # a_imp[8, i] = b[i]
    A_exp[1][0] = c[1]                                 # a_exp[2, 1] = c[2]
    A_exp[s - 1][0] = A_exp[s - 1][1]                  # a_exp[s, 1] = a_exp[s, 2]
    c[0] = Fraction(0)
    c[s - 1] = Fraction(1)
    pack = lambda f: [f.numerator, f.denominator]
    return {"s": s, "a_exp": [[pack(x) for x in r] for r in A_exp], "a_imp": [[pack(x) for x in r] for r in A_imp],
            "b": [pack(x) for x in b], "c": [pack(x) for x in c],
            "n_parsed": {k: len(v) for k, v in tabs.items()}}


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    path = os.path.join(ref, "src", "solvers", "imex_tableaus.jl")
    src = open(path, encoding="utf-8").read()
    out = {"_source": "src/solvers/imex_tableaus.jl:883-1066 (ClimaTimeSteppers.jl v0.10.0), parsed by oracle/make_kc2019_tables.py",
           "ARK437L2SA1": parse(src, "ARK437L2SA1", 7), "ARK548L2SA2": parse(src, "ARK548L2SA2", 8)}
    with open(os.path.join(HERE, "kc2019_tables.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    for k in ("ARK437L2SA1", "ARK548L2SA2"):
        print(k, out[k]["n_parsed"])


if __name__ == "__main__":
    main()
