#!/usr/bin/env python
"""One-screen summary of bench.py JSON lines:  python tools/show_bench.py [-v] file.json ..."""
import json
import sys

for f in [x for x in sys.argv[1:] if x != "-v"]:
    for line in open(f):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        r = d.get("roofline", {})
        print(f, "| %s | value %.1f e2e %.1f%s | %.2f ms/step | clk %s %s | conv frac %.3f (%.0f TF)" % (
            d["config"].get("mode", d.get("impl", "")), d["value"], d["e2e"]["value"],
            " lines %.1f" % d["e2e_text_lines"]["value"] if "e2e_text_lines" in d else "", d["ms_per_step"],
            (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("reasons"), r.get("frac", 0), r.get("achieved", 0)))
        if "stage_ms_per_step" in d:
            print("  stages", {k[:28]: round(v, 3) for k, v in d["stage_ms_per_step"].items()})
        for e in d.get("roofline_extra", []):
            print("  extra  %-30s frac %.3f  %.3f ms" % (e["kernel"], e["frac"], e["ms_per_step"]))
        if "alt_modes" in d:
            print("  alt   ", {k: "%.1f img/s, conv %.2f ms" % (v["value"], v["conv_ms_per_step"]) for k, v in d["alt_modes"].items()})
        if "parity" in d:
            print("  parity", [(i["shape"], "cls %.2e box %.2e rows %.3f" % (i["head_cls_max_abs"], i["head_bbox_max_abs"], i["oracle_rows_matched_within_1e-3"])) for i in d["parity"]["images"]])
        if "layers" in d and "-v" in sys.argv:
            for l in d["layers"]:
                print("    %-52s %8.3f ms  %7.1f alg TF" % (l["kernel"], l["ms"], l["alg_tflops"]))
