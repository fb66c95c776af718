#!/usr/bin/env python
"""Golden vectors for the result decoder: runs the REFERENCE's own resolve.py over a sample server
document and tests/tpch10noorder/dictionary.csv and stores stdin / stdout as tests/golden/resolve_*.{json,csv}.

resolve.py is Python 2.7 and this image has only Python 3 (and no lib2to3), so the source is read from /root/reference
at generation time, five Python-2-only spellings are rewritten IN MEMORY (has_key, `print >>`, indexing dict views) and
the result is executed in a subprocess; nothing of the reference is copied into the repository.  Needs /root/reference:
run here, not on the GPU box; the fixtures it writes are committed."""
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
REF = "/root/reference"

SAMPLES = {
    # the document quoted in resolve.py:8-32 (Q4's two outputs)
    "q04": {"results": {"tmp66": {".o_orderpriority__orders__o_orderpriority": [16, 40, 72, 104, 128]},
                        "tmp75": {".order_count": [311, 263, 266, 274, 36783]}},
            "timings": {"timeInMicrosecondsForFragment12": 215, "timeInMicrosecondsForFragment13": 565}},
    # dictionary-coded + plain + a code without a dictionary entry + columns of different lengths + a 4-part name
    "mixed": {"results": {"tmp0": {".n_name__nation__n_name": [72, 96, 1234]},
                          "tmp1": {".revenue": [5, 6, 7, 8]},
                          "tmp2": {".l_quantity__lineitem__l_quantity": [100, 200]},
                          "tmp3": {".a__b__c__d": [1]}},
              "timings": {}},
}


def py3_source() -> str:
    src = open(os.path.join(REF, "resolve.py")).read()
    src = re.sub(r"(\w+)\.has_key\(([^)]*)\)", r"(\2 in \1)", src)
    src = re.sub(r"print >> sys\.stderr, (.*)", r"print(\1, file=sys.stderr)", src)
    src = src.replace("res.keys()[0]", "list(res.keys())[0]").replace("res.values()[0]", "list(res.values())[0]")
    return src


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "resolve_py3.py")
        open(path, "w").write(py3_source())
        for name, doc in SAMPLES.items():
            text = json.dumps(doc)
            r = subprocess.run([sys.executable, path, os.path.join(REF, "tests", "tpch10noorder", "dictionary.csv")], input=text.encode(),
                               capture_output=True, check=True)
            open(os.path.join(out_dir, f"resolve_{name}.json"), "w").write(text)
            open(os.path.join(out_dir, f"resolve_{name}.csv"), "wb").write(r.stdout)
            print(name, r.stdout)


if __name__ == "__main__":
    main()
