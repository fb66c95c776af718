#!/usr/bin/env python
"""Prints `foreign import ccall` lines for the entry points of include/vdl_cuda.h that hs/VdlCuda.hs does not import yet
(types mapped mechanically; opaque structs become empty data declarations)."""
import os
import re

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
h = open(os.path.join(ROOT, "include", "vdl_cuda.h")).read()
h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
hs = open(os.path.join(ROOT, "hs", "VdlCuda.hs")).read()
have = set(re.findall(r'foreign import ccall \w+ "(vdl_[a-z0-9_]+)"', hs))

OPAQUE = {"vdl_ctx": "VdlCtx", "vdl_plan": "VdlPlan", "vdl_fused": "VdlFused", "vdl_probe": "VdlProbe", "vdl_fused_desc": "VdlFusedDesc",
          "vdl_probe_desc": "VdlProbeDesc", "vdl_map_desc": "VdlMapDesc", "vdl_fold_spec": "VdlFoldSpec", "vdl_comm_plan": "VdlCommPlan", "vdl_comm": "VdlComm"}
BASE = {"int": "CInt", "int32_t": "Int32", "int64_t": "Int64", "uint64_t": "Word64", "float": "CFloat", "vdl_vec": "VdlVec", "void": "()", "char": "CChar"}


def hs_type(c: str) -> str:
    c = c.replace("const", " ").strip()
    stars = c.count("*")
    base = c.replace("*", " ").split()[0]
    if base == "char" and stars >= 1:
        t, stars = "CString", stars - 1
    else:
        t = OPAQUE.get(base) or BASE[base]
    for _ in range(stars):
        t = f"Ptr {t}" if " " not in t else f"Ptr ({t})"
    return t


for m in re.finditer(r"^([A-Za-z_][\w \*]*?)\b(vdl_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", h, flags=re.M | re.S):
    ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
    if name in have:
        continue
    params = []
    if args and args != "void":
        for a in args.split(","):
            a = a.strip()
            a = re.sub(r"\b[a-z_][a-z0-9_]*$", "", a).strip() if not a.endswith("*") else a     # drop the parameter name
            params.append(hs_type(a))
    r = hs_type(ret)
    r = f"IO {r}" if " " not in r else f"IO ({r})"
    print(f'foreign import ccall safe "{name}" c_{name} :: ' + " -> ".join(params + [r]))
