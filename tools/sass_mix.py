#!/usr/bin/env python3
"""Opcode / pipe mix of a SASS line range: sass_mix.py file first_line last_line"""
import collections
import re
import sys
ALU = {"IADD3", "LOP3", "SEL", "ISETP", "LEA", "MOV", "SHF", "PRMT", "VIADD", "IABS", "IMNMX", "VIMNMX", "ISCADD", "PLOP3", "P2R", "R2P", "FSEL", "CS2R", "UMOV"}
FMA = {"IMAD", "HFMA2", "FFMA", "FMUL", "FADD"}
lines = open(sys.argv[1]).read().splitlines()
a, b = int(sys.argv[2]), int(sys.argv[3])
ops = collections.Counter()
full = collections.Counter()
for ln in lines[a - 1:b]:
    toks = ln.split()
    op = toks[1] if toks[1].startswith("@") is False else toks[2]
    if toks[1].startswith("@"):
        op = toks[2]
    base = op.split(".")[0].rstrip(";")
    ops[base] += 1
    full[op.rstrip(";")] += 1
tot = sum(ops.values())
alu = sum(v for k, v in ops.items() if k in ALU)
fma = sum(v for k, v in ops.items() if k in FMA)
print("total %d  alu %d (%.0f%%)  fma %d (%.0f%%)  other %d" % (tot, alu, 100 * alu / tot, fma, 100 * fma / tot, tot - alu - fma))
print(", ".join("%s %d" % kv for kv in full.most_common(18)))
