#!/usr/bin/env python
"""Static per-iteration instruction mix of a kernel's main loop, from `cuobjdump -sass`.

    python scripts/sass_loop_profile.py <object.o> <regex on the mangled kernel name> [...]

The main loop is taken to be the LARGEST backward branch (highest instruction count between the branch target and
the branch).  Instructions are grouped by issue pipe as on sm_100 (alu: integer / logic / FMNMX / select /
predicate-setting; fma: FADD / FMUL / FFMA / IMAD; lsu: LDS / LDG / STG / LDGSTS; ...).  This is a static count --
predicated-off and branched-over instructions are included -- so it is an upper bound of the dynamic mix; it is
what the notes in profiles/cw_select.md section 4 are computed from.
"""
import collections
import re
import subprocess
import sys

ALU = {"FMNMX", "FMNMX3", "ISETP", "LEA", "FSEL", "LOP3", "IADD3", "SEL", "VIADD", "FSETP", "SHF", "PLOP3", "IADD",
       "VIMNMX", "VIMNMX3", "MOV", "IABS", "PRMT", "UISETP", "ULOP3", "UIADD3", "ULEA", "USEL", "UMOV", "USHF", "UPLOP3",
       "UIMAD", "UFLO", "FLO", "POPC", "BREV", "P2R", "R2P", "R2UR", "CS2R", "S2R", "S2UR", "VOTE", "VOTEU", "REDUX"}
FMA = {"FADD", "FMUL", "FFMA", "IMAD", "HFMA2", "FCHK"}
LSU = {"LDS", "STS", "LDG", "STG", "LDGSTS", "LD", "ST", "LDL", "STL", "ATOM", "ATOMG", "RED", "LDSM", "LDGDEPBAR", "DEPBAR"}
CONST = {"LDC", "LDCU", "ULDC"}
CTRL = {"BRA", "EXIT", "BSSY", "BSYNC", "NOP", "WARPSYNC", "BAR", "CALL", "RET", "YIELD", "BPT", "ERRBAR", "MEMBAR"}
XU = {"MUFU", "F2I", "I2F", "F2F", "I2FP", "F2FP"}


def pipe_of(op):
    for name, grp in (("alu", ALU), ("fma", FMA), ("lsu", LSU), ("const", CONST), ("ctrl", CTRL), ("xu", XU)):
        if op in grp:
            return name
    return "other"


def kernels(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        yield f.split("\n")[0].strip(), f


def profile(body):
    ins = []          # (address, opcode, branch target or None)
    for m in re.finditer(r"^\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z][A-Z0-9_]*)([^;]*);", body, re.M):
        addr, op, rest = int(m.group(1), 16), m.group(2), m.group(3)
        tgt = None
        if op == "BRA":
            t = re.search(r"0x([0-9a-f]+)", rest)
            tgt = int(t.group(1), 16) if t else None
        ins.append((addr, op, tgt))
    best = None
    for i, (addr, op, tgt) in enumerate(ins):
        if op == "BRA" and tgt is not None and tgt < addr:
            span = [x for x in ins if tgt <= x[0] <= addr]
            if best is None or len(span) > len(best):
                best = span
    whole = collections.Counter(op for _, op, _ in ins)
    loop = collections.Counter(op for _, op, _ in (best or []))
    return whole, loop


def main(argv):
    obj, pats = argv[0], argv[1:]
    for name, body in kernels(obj):
        if not any(re.search(p, name) for p in pats):
            continue
        whole, loop = profile(body)
        pipes = collections.Counter()
        for op, c in loop.items():
            pipes[pipe_of(op)] += c
        print(f"{name}\n  whole kernel: {sum(whole.values())} instructions; main loop: {sum(loop.values())}"
              f"\n  loop by pipe: {dict(pipes)}\n  loop opcodes: {dict(loop.most_common(16))}")
        other = {op: c for op, c in loop.items() if pipe_of(op) == "other"}
        if other:
            print(f"  unclassified: {other}")


if __name__ == "__main__":
    main(sys.argv[1:])
