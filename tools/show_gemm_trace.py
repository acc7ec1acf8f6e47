"""Condense `bench.py --gemm-trace` output (clock stamps of CTA 0 of each tcgen05 GEMM launch of one eager step)."""
import json, sys

for path in sys.argv[1:]:
    d = None
    for line in open(path):
        line = line.strip()
        if line.startswith("{") and "gemm_trace" in line:
            d = json.loads(line)["gemm_trace"]
    print(path)
    if d is None:
        print("  no gemm_trace line"); continue
    for g in d:
        kb = g["kb"]
        n = len(kb)
        tma = [k[1] - k[0] for k in kb]                    # TMA issue -> landed (seen by the transform warps)
        xf = [k[2] - k[1] for k in kb]                     # transform
        wait = [k[6] - k[5] for k in kb]                   # MMA thread waiting for the transform
        iss = [k[4] - k[3] for k in kb]                    # MMA issue
        period = [(kb[i + 1][2] - kb[i][2]) for i in range(n - 1)] or [0]
        steady = period[len(period) // 3:] or period
        med = lambda v: sorted(v)[len(v) // 2]
        print(f"  mode {'TN' if g['mode'] else 'NT'} BN {g['BN']} stages {g['stages']} k-blocks {g['num_kb']}: "
              f"period med {med(steady)} clk (TMA {med(tma)}, transform {med(xf)}, MMA wait {med(wait)}, issue {med(iss)}); "
              f"main loop end {kb[-1][4]}, epilogue {g['epi_start']}..{g['epi_end']}")
