#!/usr/bin/env python
"""Condensed view of a bench.py JSON line and (optionally) a --trace JSON: python tools/show_bench.py bench.json [trace.json]"""
import json
import sys


def main():
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.0f mol/s  %.4f ms/step  passes %s  launches/step %s" % (
        d["value"], d["ms_per_step"], d.get("impl_detail", {}).get("timed_passes_ms_per_step"), d.get("gpu_launches_per_step")))
    print({k: round(d[k]["value"]) for k in ("e2e", "e2e_full_copy", "e2e_packed") if k in d})
    r = d.get("roofline")
    if r:
        print("roofline", r["kernel"], "frac %.4f" % r["frac"], "3xtf32 %s" % r.get("frac_of_3xtf32_ceiling"), "avg us %.1f" % r["avg_launch_us"],
              r.get("step_vs_rooflines"))
        for k, v in list(r["kernels"].items())[:12]:
            print("   %-26s %.4f ms  x%.0f" % (k, v["ms_per_step"], v["launches_per_step"]))
    if d.get("cpu_baseline"):
        c = d["cpu_baseline"]
        print("cpu", round(c["value"], 1), c["kind"], c["cores"])
    for k, v in d.get("configs", {}).items():
        if "error" in v:
            print(k, v)
            continue
        rr = v.get("roofline") or {}
        print(k, "%.0f mol/s %.4f ms launches %s | top %s frac %.4f 3xtf32 %s | %s" % (
            v["value"], v["ms_per_step"], v["gpu_launches_per_step"], rr.get("kernel"), rr.get("frac", 0),
            rr.get("frac_of_3xtf32_ceiling"), rr.get("step_vs_rooflines")))
        for kk, vv in list(rr.get("kernels", {}).items())[:6]:
            print("      %-26s %.4f ms  x%.0f" % (kk, vv["ms_per_step"], vv["launches_per_step"]))
    if len(sys.argv) > 2:
        t = json.load(open(sys.argv[2]))
        print("trace span %.1f us busy %.1f us" % (t["span_us_per_step"], t["busy_us_per_step"]))
        for e in t["last_step_sequence"]:
            print("  %6.1f %6.1f %s" % (e["us"], e["gap_us"], e["name"][:48]))


if __name__ == "__main__":
    main()
