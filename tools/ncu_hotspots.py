#!/usr/bin/env python
"""Per-source-line hot spots of one kernel launch from an Nsight Compute report captured with --import-source on.

    ncu -i gpurun_out/<tag>_prof.ncu-rep --page source --csv --print-source cuda,sass \
        --kernel-name regex:<kernel> --launch-skip <n> --launch-count 1 > /tmp/k.csv
    python tools/ncu_hotspots.py /tmp/k.csv [top_n]

Prints, for the top_n CUDA source lines by warp-state samples: share of the kernel's samples, share of its executed warp
instructions, the three most frequent stall reasons and the source text (profiles/r01s2_ncu_hotspots.txt was made this way).
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    header, cur_file = None, "?"
    per = collections.defaultdict(lambda: [0, 0, "", collections.Counter()])
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1]
            continue
        if "# Samples" in r:
            header = r
            si, ii = header.index("# Samples"), header.index("Instructions Executed")
            stall_cols = [(i, c) for i, c in enumerate(header) if c.startswith("stall_") and "Not Issued" not in c]
            continue
        if header is None or len(r) < len(header) - 5 or r[0] == "":      # SASS rows repeat their source line's totals
            continue
        try:
            samples, instr = int(r[si]), int(r[ii])
        except ValueError:
            continue
        p = per[(cur_file.split("/")[-1], int(r[0]))]
        p[0] += samples
        p[1] += instr
        p[2] = r[1][:100]
        for i, c in stall_cols:
            try:
                p[3][c] += int(r[i])
            except ValueError:
                pass
    tot = sum(p[0] for p in per.values()) or 1
    toti = sum(p[1] for p in per.values()) or 1
    print("total samples", tot, "instr", toti)
    for k, p in sorted(per.items(), key=lambda kv: -kv[1][0])[:top_n]:
        top = ", ".join(f"{c[6:]}:{n}" for c, n in p[3].most_common(3))
        print(f"{k[0]}:{k[1]:>4} {100 * p[0] / tot:5.1f}% smp {100 * p[1] / toti:5.1f}% ins | {top} | {p[2]}")


if __name__ == "__main__":
    main()
