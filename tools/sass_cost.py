#!/usr/bin/env python3
"""Issue-time estimate of a kernel from its SASS.  Costs fitted on B200 to five measured integer kernels
(tools/microbench/poseidon_bench.cu: Poseidon-GL old and new forms, its S-box and MDS chains alone, Poseidon2-BB;
all within 2%): the integer path sustains ~0.70 warp-instructions per cycle per scheduler (1.42 cycles each) whatever
the opcode (IADD3, IADD3.X, LEA, SEL, LOP3, IMAD.MOV/IADD/X/SHL, 32-bit IMAD), and IMAD.WIDE / IMAD.HI cost 2.6 in
mixed code (a back-to-back stream of them is pipe-limited at 4.1, tools/microbench/imad_probe.cu).
usage: sass_cost.py <object-or-cubin> <mangled kernel substring>  -> per-loop and static totals"""
import re
import subprocess
import sys

WIDE = 2.59
OTHER = 1.42
FREE = ("NOP",)


def cost(op):
    base = op.split(".")
    if base[0] == "IMAD" and len(base) > 1 and base[1] in ("WIDE", "HI"):
        return WIDE
    if base[0] in FREE:
        return 0.0
    return OTHER


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, funcs = None, {}
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
        if m and cur:
            toks = m.group(2).split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            funcs[cur].append((int(m.group(1), 16), op, m.group(2)))
    for name, ins in funcs.items():
        if pat not in name:
            continue
        addr_idx = {a: i for i, (a, _, _) in enumerate(ins)}
        print("== %s: %d instructions, static cost %.0f" % (name[:80], len(ins), sum(cost(o) for _, o, _ in ins)))
        # backward branches = loops
        for i, (a, op, text) in enumerate(ins):
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", text)
                if m and int(m.group(1), 16) in addr_idx and int(m.group(1), 16) < a:
                    j = addr_idx[int(m.group(1), 16)]
                    body = ins[j:i + 1]
                    wide = sum(1 for _, o, _ in body if o.startswith("IMAD.WIDE") or o.startswith("IMAD.HI"))
                    print("   loop %05x..%05x: %4d instr, %3d wide/hi, cost %.0f cycles" % (ins[j][0], a, len(body), wide, sum(cost(o) for _, o, _ in body)))


if __name__ == "__main__":
    main()
