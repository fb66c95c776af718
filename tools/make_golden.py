#!/usr/bin/env python
"""Extract the golden vectors the reference holds for this path into tests/golden/.

The only executable golden vector in orm011/mplan2vdl is the Q6 plan excerpt printed in its
README (README.md:40-52: statements 1-9, an ellipsis, statements 40-42).  Run in the build
container (needs /root/reference): python tools/make_golden.py
"""
import os
import re

ref = "/root/reference/README.md"
dst = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "readme_q06_lines.txt")
lines = []
grab = False
for ln in open(ref).read().splitlines():
    if ln.startswith("$ ./tpchrun"):
        grab = True
        continue
    if grab and (re.match(r"^\d+,", ln) or ln.strip() == "..."):
        lines.append(ln.rstrip())
    elif grab and lines and ln.startswith("```"):
        break
with open(dst, "w") as f:
    f.write("\n".join(lines) + "\n")
print(f"wrote {dst}: {len(lines)} lines")
