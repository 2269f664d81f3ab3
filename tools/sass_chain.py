#!/usr/bin/env python3
"""In-order issue model of a straight-line SASS block (one warp, one instruction per cycle at best): how many cycles one
warp needs for the block when every result takes its pipe latency — the dependency-chain floor of the rollout loop.
Reads `addr exec text` lines (tools/ncu_hot-style dump) or cuobjdump lines on stdin;  --lat-fma 4 --lat-mufu 18 ...
"""
import re
import sys

LAT = {"fma": 4, "alu": 4, "mufu": 18, "lds": 30, "ldcu": 25, "f2f": 10, "dadd": 8}
FMA = ("FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "FFMA2", "FMUL2", "FADD2")
ALU = ("FMNMX", "LOP3", "IADD3", "SHF", "SEL", "FSEL", "ISETP", "FSETP", "PRMT", "MOV", "LEA", "VIADD", "PLOP3")


def kind(op):
    b = op.split(".")[0]
    if b in FMA: return "fma"
    if b in ALU: return "alu"
    if b == "MUFU": return "mufu"
    if b in ("LDS", "LDC"): return "lds"
    if b == "LDCU": return "ldcu"
    if b in ("F2F", "I2F", "F2I", "I2FP"): return "f2f"
    if b in ("DADD", "DFMA", "DMUL"): return "dadd"
    return "alu"


def regs(tok):
    return re.findall(r"\b(?:UR|R|P|UP)\d+\b", tok)


def main():
    lines = [l.strip() for l in sys.stdin if l.strip()]
    ready = {}
    t = 0
    last_kind = None
    n = 0
    pipe_free = {"fma": 0, "alu": 0, "mufu": 0}
    rt = {"fma": float(sys.argv[sys.argv.index("--rt-fma") + 1]) if "--rt-fma" in sys.argv else 1.0,
          "alu": float(sys.argv[sys.argv.index("--rt-alu") + 1]) if "--rt-alu" in sys.argv else 1.0, "mufu": 8.0}
    for l in lines:
        m = re.match(r"^\S+\s+\d+\s+(.*)$", l)
        text = m.group(1) if m else l
        text = re.sub(r"^@!?U?P\d+\s+", "", text)
        parts = text.split(None, 1)
        op = parts[0]
        if op in ("BRA", "BSSY", "BSYNC", "NOP"):
            continue
        ops = parts[1].split(",") if len(parts) > 1 else []
        k = kind(op)
        dst = regs(ops[0]) if ops else []
        src = [r for o in ops[1:] for r in regs(o)]
        wide = ".128" in op or ".WIDE" in op or ".64" in op
        dsts = list(dst)
        if dst and wide and dst[0].startswith("R"):
            base = int(dst[0][1:])
            cnt = 4 if ".128" in op else 2
            dsts = [f"R{base+i}" for i in range(cnt)]
        if dst and wide and dst[0].startswith("UR"):
            base = int(dst[0][2:])
            cnt = 4 if ".128" in op else 2
            dsts = [f"UR{base+i}" for i in range(cnt)]
        start = t + 1
        for r in src:
            if r in ready:
                start = max(start, ready[r] + (1 if False else 0))
        pk = k if k in pipe_free else None
        if pk:
            start = max(start, pipe_free[pk])
            pipe_free[pk] = start + rt[pk]
        t = start
        for r in dsts:
            ready[r] = t + LAT[k]
        n += 1
    print(f"{n} instructions, single-warp in-order time {t} cycles, {t/n:.2f} cycles/instr")


if __name__ == "__main__":
    main()
