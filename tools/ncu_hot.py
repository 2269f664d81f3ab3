#!/usr/bin/env python3
"""Per-instruction view of an .ncu-rep source page (read here, no GPU): the SASS lines of the first profiled kernel whose
name contains a pattern, with executed counts and the dominant stall reasons, restricted to the hottest region.

  python tools/ncu_hot.py gpurun_out/prof.ncu-rep [kernel-substring] [--min-exec N] [--top N]
"""
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
    min_exec = int(sys.argv[sys.argv.index("--min-exec") + 1]) if "--min-exec" in sys.argv else 0
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 0
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    blocks = out.split('"Kernel Name",')
    for b in blocks[1:]:
        name, _, rest = b.partition("\n")
        if pat not in name:
            continue
        rows = list(csv.reader(io.StringIO(rest)))
        hdr = rows[0]
        col = {h: i for i, h in enumerate(hdr)}
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        data = [r for r in rows[1:] if len(r) == len(hdr)]
        tot = sum(int(r[col["# Samples"]] or 0) for r in data)
        print("==", name.strip()[:120], "samples", tot)
        if top:
            ranked = sorted(data, key=lambda r: -int(r[col["# Samples"]] or 0))[:top]
            keep = set(id(r) for r in ranked)
        for r in data:
            ex = int(r[col["Instructions Executed"]] or 0)
            if ex < min_exec:
                continue
            if top and id(r) not in keep:
                continue
            smp = int(r[col["# Samples"]] or 0)
            st = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
            sts = " ".join(f"{n}:{v}" for v, n in st if v)
            print(f"{r[col['Address']][-5:]} {ex:8d} {smp:6d} {100.0*smp/max(tot,1):5.2f}%  {r[col['Source']].strip():70s} {sts}")
        break


if __name__ == "__main__":
    main()
