#!/usr/bin/env python
"""Regenerate the results table of DESIGN.md section 6 from the committed bench lines (profiles/r2_bench_*.json)."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
order = ["cfg2", "cfg2_steps200", "cfg3", "cfg3_steps200", "cfg4", "cfg5", "4gpu", "8gpu_box_1gpu", "8gpu", "8gpu_cfg3", "8gpu_cfg5", "reference_arm"]
files = [os.path.join(ROOT, "profiles", "r2_bench_%s.json" % k) for k in order]
files = [f for f in files if os.path.exists(f)]
table = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "results_table.py")] + files, capture_output=True, text=True).stdout
p = os.path.join(ROOT, "DESIGN.md")
s = open(p).read()
s = re.sub(r"<!-- results:begin -->.*?<!-- results:end -->", "<!-- results:begin -->\n" + table.strip() + "\n<!-- results:end -->", s, flags=re.S)
open(p, "w").write(s)
print(table)
