#!/usr/bin/env python
"""prints the key figures of a bench.py JSON line (file argument or stdin)"""
import json
import sys

txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
d = json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
r = d.get("roofline", {})
print("ms/step %.4f  value %.0f  e2e %.0f  dominant %s frac %.3f  pipeline_alg_frac %.3f  launches %s" % (
    d["ms_per_step"], d["value"], d["e2e"]["value"], r.get("kernel"), r.get("frac", 0), r.get("pipeline_alg_frac", 0),
    d.get("gpu_launches")))
print(r.get("kernels_ms_per_step"))
for k, v in (d.get("extra") or {}).items():
    print("extra.%s: %s" % (k, {kk: vv for kk, vv in v.items() if kk in ("value", "ms_per_slide", "ms_per_batch", "ms_per_pass",
                                                                          "ms_per_step", "ms_per_tile", "verified", "error",
                                                                          "plan_graph_ms_per_tile", "plan_stream_ms_per_tile")}))
