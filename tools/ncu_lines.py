"""Per-source-line shares of a kernel's executed warp instructions and stall samples, from an ncu report captured with
--set full --import-source on (and -lineinfo in the build).

    ncu -i gpurun_out/prof.ncu-rep --page source --csv --print-source sass,cuda > /tmp/src.csv
    python tools/ncu_lines.py /tmp/src.csv [min_percent]

Prints the totals, the share of every source FILE, then every line with at least min_percent (default 0.4) of the
instructions or of the samples.  Out-of-line subroutines (IEEE division, libm slow paths) are attributed by ncu to the
line of the first inlined intrinsic of the caller; read the SASS view of that line to tell them apart."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
    cur = None
    inst, samp, text = collections.Counter(), collections.Counter(), {}
    for r in csv.reader(open(path)):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) < 8 or r[0] in ("Line No", ""):
            continue  # header, or a SASS row (those have an empty line number)
        try:
            line, n_inst, n_samp = int(r[0]), int(r[7]), int(r[6])
        except ValueError:
            continue
        inst[(cur, line)] += n_inst
        samp[(cur, line)] += n_samp
        text[(cur, line)] = r[1].strip()[:90]
    ti, ts = sum(inst.values()), max(sum(samp.values()), 1)
    print(f"warp instructions {ti}   stall samples {ts}")
    by_file_i, by_file_s = collections.Counter(), collections.Counter()
    for k, v in inst.items():
        by_file_i[k[0]] += v
        by_file_s[k[0]] += samp[k]
    print("file: % instructions, % samples")
    for f, v in by_file_i.most_common():
        print(f"  {f:28s} {100 * v / ti:5.1f} {100 * by_file_s[f] / ts:5.1f}")
    print("line: % instructions, % samples")
    for k in sorted(inst):
        pi, ps = 100 * inst[k] / ti, 100 * samp[k] / ts
        if pi >= thr or ps >= thr:
            print(f"  {k[0]}:{k[1]:<5d} {pi:5.2f} {ps:5.2f}  {text[k]}")


if __name__ == "__main__":
    main()
