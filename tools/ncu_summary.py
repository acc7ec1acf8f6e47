#!/usr/bin/env python
"""Condense an Nsight Compute report into what profiles/ keeps: a selected-column CSV of every captured launch and
profiles/traffic.json (DRAM read + write bytes per launch and kernel: bench.py's roofline.traffic).

    ncu -i gpurun_out/<tag>_prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv profiles/<tag>_ncu_full_summary.csv [profiles/traffic.json]

Round 2: the forward projections live inside layer_fwd_fused_kernel; the launch order of gemm_tc_kernel inside one eager
step of bench.py is dH2 (nt), dW2 (tn), dW1 (tn).  mm_tile_kernel<..> instantiations are tagged nn / nt / tn / tt.
"""
import collections
import csv
import json
import re
import sys

KEEP = ["ID", "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
TAGS = {"agg_bwd_tile_kernel": "agg_bwd_kernel", "agg_fwd_tile_kernel": "agg_fwd_kernel",
        "bn_bwd_partial_vec_kernel": "bn_bwd_partial_kernel", "bn_act_fwd_vec_kernel": "bn_act_fwd_kernel",
        "bn_act_bwd_vec_kernel": "bn_act_bwd_kernel"}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h, units = rows[0], rows[1]
    cols = [c for c in KEEP if c in h]
    idx = [h.index(c) for c in cols]
    with open(sys.argv[2], "w", newline="") as f:
        w = csv.writer(f)
        for r in rows:
            w.writerow([r[i] for i in idx])
    if len(sys.argv) < 4:
        return
    iname, ir, iw = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    mr, mw = mult[units[ir]], mult[units[iw]]
    by = collections.defaultdict(list)
    for r in rows[2:]:
        k = re.match(r"(?:void )?(?:eagcn::)?(?:tc::|fz::)?(\w+)", r[iname]).group(1)
        if k == "layer_fwd_fused_kernel":
            k = "layer_fwd_fused"
        if k == "mm_tile_kernel":
            m = re.search(r"mm_tile_kernel<\(bool\)(\d), \(bool\)(\d)>|mm_tile_kernel<(true|false), (true|false)>", r[iname])
            if m:
                a, b = (m.group(1), m.group(2)) if m.group(1) is not None else (str(int(m.group(3) == "true")), str(int(m.group(4) == "true")))
                k = "mm_tile_" + ("n" if a == "1" else "t") + ("t" if b == "1" else "n")
        by[TAGS.get(k, k)].append(float(r[ir]) * mr + float(r[iw]) * mw)
    out = {"_source": f"{sys.argv[2]} (ncu --set full --clock-control none, eager steps, cold caches): dram__bytes_read.sum + "
                      "dram__bytes_write.sum per launch, averaged over the captured launches of each kernel"}
    for k, v in by.items():
        out[k] = sum(v) / len(v)
    json.dump(out, open(sys.argv[3], "w"), indent=1)


if __name__ == "__main__":
    main()
