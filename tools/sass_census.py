#!/usr/bin/env python3
"""SASS census: per kernel of an object file, registers/spills (from the .ptxas.log next to it) and, for every loop
(backward branch), the instruction count by pipe class.  Used for the per-step instruction budget in DESIGN.md.

  python tools/sass_census.py mpc_rs_b200/csrc/mppi_ws_NL.o [name-substring] [--loops] [--min N]
"""
import re
import subprocess
import sys
from collections import Counter

FMA = ("FFMA", "FMUL", "FADD", "IMAD", "FFMA2", "FMUL2", "FADD2", "HFMA2", "IMUL")
ALU = ("FMNMX", "LOP3", "IADD3", "IADD", "SHF", "SEL", "FSEL", "ISETP", "FSETP", "PRMT", "MOV", "LEA", "PLOP3", "FMNMX3", "VIADD",
       "IABS", "FCHK", "SGXT", "BMSK", "VIMNMX", "IMNMX", "CS2R", "P2R", "R2P", "FSET", "FLO", "POPC", "BREV", "SHL", "SHR")
XU = ("MUFU", "I2F", "F2I", "F2F", "I2FP", "F2FP", "FRND")
FP64 = ("DADD", "DMUL", "DFMA", "DSETP", "DMNMX")
MEM = ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDC", "LDCU", "ATOM", "ATOMS", "RED", "LDSM", "LDL", "STL", "ATOMG", "SYNCS", "MEMBAR", "ERRBAR", "CCTL")
CTRL = ("BRA", "BSSY", "BSYNC", "EXIT", "RET", "CALL", "BAR", "WARPSYNC", "NOP", "YIELD", "BREAK", "JMP", "BRX", "DEPBAR", "NANOSLEEP", "BMOV", "ENDCOLLECTIVE")
WARP = ("SHFL", "VOTE", "VOTEU", "REDUX", "MATCH", "S2R", "S2UR", "R2UR", "UMOV", "ULDC", "UIADD3", "ULOP3", "USHF", "UISETP", "UIMAD", "UMOV", "ULEA", "UFLO", "UPOPC", "USEL", "UPRMT", "UF2FP", "UPLOP3", "R2UR", "UFMUL", "UFADD", "UFFMA", "UI2FP", "UF2F", "USGXT", "UBMSK", "UFSETP", "UFSEL", "UFMNMX", "UVIADD", "UVIMNMX", "UIABS")


def klass(op):
    base = op.split(".")[0]
    if base.startswith("U") and base not in ("UTMALDG", "UBLKCP"):
        return "uniform"
    for name, group in (("fma", FMA), ("alu", ALU), ("xu", XU), ("fp64", FP64), ("mem", MEM), ("ctrl", CTRL), ("warp", WARP)):
        if base in group:
            return name
    return "other:" + base


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, body = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if cur:
                yield cur, body
            cur, body = m.group(1), []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            addr = int(m.group(1), 16)
            text = m.group(2).strip()
            pred = ""
            pm = re.match(r"(@!?U?P\d+)\s+(.*)", text)
            if pm:
                pred, text = pm.group(1), pm.group(2)
            op = text.split()[0]
            body.append((addr, op, text, pred))
    if cur:
        yield cur, body


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    obj = args[0]
    sub = args[1] if len(args) > 1 else ""
    show_loops = "--loops" in sys.argv
    min_len = 20
    if "--min" in sys.argv:
        min_len = int(sys.argv[sys.argv.index("--min") + 1])
    for name, body in functions(obj):
        if sub not in name:
            continue
        print(f"== {name}: {len(body)} instructions")
        if not show_loops:
            continue
        index = {a: i for i, (a, _, _, _) in enumerate(body)}
        loops = []
        for i, (a, op, text, _) in enumerate(body):
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", text)
                if m:
                    tgt = int(m.group(1), 16)
                    if tgt <= a and tgt in index:
                        loops.append((index[tgt], i))
        for lo, hi in sorted(loops):
            n = hi - lo + 1
            if n < min_len:
                continue
            inner = [(l, h) for (l, h) in loops if l >= lo and h <= hi and (l, h) != (lo, hi)]
            cnt = Counter(klass(op) for _, op, _, _ in body[lo:hi + 1])
            ops = Counter(op.split(".")[0] for _, op, _, _ in body[lo:hi + 1])
            tag = "innermost" if not inner else f"contains {len(inner)} loops"
            print(f"  loop @{body[lo][0]:#06x}-{body[hi][0]:#06x}: {n} instr ({tag})  " +
                  "  ".join(f"{k}={v}" for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])))
            print("      " + " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    main()
