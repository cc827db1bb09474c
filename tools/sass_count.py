#!/usr/bin/env python
"""Per-category SASS instruction count of a kernel's hot loop.

    python tools/sass_count.py [--obj sdetools_b200/build/ensemble_model_1.o] [--kernel REGEX] [--steps-per-iter 4] [--json OUT]

The hot loop is taken to be the innermost backward branch that spans the most instructions and contains FP64
arithmetic.  Every instruction between the branch target and the branch is classified by opcode and, for the
integer ones, by what it feeds:
  philox   IMAD.WIDE.U32 / IMAD.HI.U32 / IMAD x constant and the LOP3 that mix the round keys (UR operands)
  fp64     DFMA DMUL DADD DSETP (the FP64 pipe), MUFU.RSQ64H counted apart
  loads    LDG LDS LDC and their address arithmetic (IMAD.WIDE to an address, LEA)
  control  BRA BSSY BSYNC ISETP on the loop counter, VIADD of the counter
  other    everything else (quantile index bit-field work, sign handling, moves)
Counts are per loop iteration and per path-step (iteration / --steps-per-iter).  The script reads cuobjdump
output, needs no GPU, and is what bench.py cites for the issue-slot bound.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTR = re.compile(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);")


def disassemble(obj: str):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), [])
            continue
        m = INSTR.match(line)
        if m and cur is not None:
            cur.append((int(m.group(1), 16), m.group(2), m.group(3).strip(), line))
    return kernels


def hot_loop(ins):
    """(start index, end index) of the innermost backward branch with the most FP64 arithmetic in its body."""
    addr_to_idx = {a: i for i, (a, *_rest) in enumerate(ins)}
    loops = []
    for i, (a, op, args, _) in enumerate(ins):
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", args)
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr_to_idx:
                j = addr_to_idx[tgt]
                nfp = sum(1 for x in ins[j:i + 1] if x[1].startswith(("DFMA", "DMUL", "DADD")))
                if nfp >= 4:
                    loops.append((j, i, nfp))
    inner = [l for l in loops if not any(o is not l and o[0] >= l[0] and o[1] <= l[1] for o in loops)]
    if not inner:
        raise SystemExit("no loop with FP64 arithmetic found")
    j, i, _ = max(inner, key=lambda l: l[2])
    return j, i


def classify(op: str, args: str) -> str:
    if op.startswith(("DFMA", "DMUL", "DADD", "DSETP")):
        return "fp64"
    if op.startswith("MUFU"):
        return "mufu"
    if op.startswith(("LDG", "LDS", "LDC", "LD.", "LDL", "STG", "STS", "STL", "ST.")):
        return "loads"
    if op.startswith(("BRA", "BSSY", "BSYNC", "CALL", "RET", "EXIT", "WARPSYNC", "NOP")):
        return "control"
    if op.startswith(("IMAD.WIDE.U32", "IMAD.HI.U32")) and re.search(r"-?0x(326172a9|2daee0ad|cd9e8d57|d2511f53)", args):
        return "philox"
    if op == "IMAD" and re.search(r"-?0x(326172a9|2daee0ad)", args):
        return "philox"
    if op.startswith("LOP3") and "UR" in args and "0x96" in args:
        return "philox"
    if op.startswith(("IMAD.MOV", "MOV", "IMAD.U32", "HFMA2", "UMOV", "S2UR", "CS2R")):
        return "moves"
    if op.startswith(("FSEL", "SEL", "VIMNMX", "FMNMX")):
        return "select"
    return "integer"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--obj", default=os.path.join(ROOT, "sdetools_b200", "build", "ensemble_model_1.o"))
    ap.add_argument("--kernel", default=r"k_terminalINS_5ModelILi1EEELi1ELi1ELb1", help="regex on the mangled kernel name")
    ap.add_argument("--steps-per-iter", type=int, default=4)
    ap.add_argument("--json", default=None)
    ap.add_argument("--listing", action="store_true", help="print the loop body")
    ap.add_argument("--loop", default="", help="START:END addresses (hex) of the loop when the heuristic picks another one, e.g. 0xc00:0x1330")
    args = ap.parse_args()
    kernels = disassemble(args.obj)
    names = [k for k in kernels if re.search(args.kernel, k)]
    if not names:
        raise SystemExit(f"no kernel matches {args.kernel}; have e.g. {list(kernels)[:3]}")
    name = names[0]
    ins = kernels[name]
    j, i = hot_loop(ins)
    if args.loop:
        lo, hi = (int(v, 16) for v in args.loop.split(":"))
        j = next(k for k, x in enumerate(ins) if x[0] == lo)
        i = next(k for k, x in enumerate(ins) if x[0] == hi)
    body = ins[j:i + 1]
    cats = {}
    for a, op, ar, _ in body:
        cats[classify(op, ar)] = cats.get(classify(op, ar), 0) + 1
    total = len(body)
    nfma = sum(1 for _, op, _, _ in body if op.startswith("DFMA"))
    nadd = sum(1 for _, op, _, _ in body if op.startswith(("DADD", "DMUL")))
    res = {"kernel": name, "loop": [hex(ins[j][0]), hex(ins[i][0])], "steps_per_iteration": args.steps_per_iter,
           "instructions_per_iteration": total, "instructions_per_path_step": total / args.steps_per_iter,
           "per_path_step": {k: v / args.steps_per_iter for k, v in sorted(cats.items())},
           "fp64_per_path_step": cats.get("fp64", 0) / args.steps_per_iter,
           "dfma_per_path_step": nfma / args.steps_per_iter, "dadd_dmul_per_path_step": nadd / args.steps_per_iter,
           "fp64_flops_per_path_step": (2 * nfma + nadd) / args.steps_per_iter}
    if args.listing:
        for a, op, ar, line in body:
            print(f"{classify(op, ar):8s} {line.strip()}")
    print(json.dumps(res, indent=1))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
