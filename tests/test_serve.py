"""The HTTP face of the executor (mplan2vdl_b200/serve.py): the wire the reference pipeline already speaks
(eval_query.sh:18-26 POSTs `text/vdl`, resolve.py:8-32 reads the JSON).  The protocol is tested here with the CPU oracle standing
in for the GPU backend (test infrastructure); tests/test_zz_serve_gpu.py runs real programs through libvdl_cuda behind the same handler."""
import json
import threading
import urllib.error
import urllib.request
from http.server import HTTPServer

import numpy as np
import pytest

from mplan2vdl_b200 import resolve, serve, synth, tpch
from util import host_columns, plan_text, run_oracle


class OracleBackend:
    def __init__(self, catalog, sf=0.01):
        self.cat, self.sf = catalog, sf

    def run(self, text):
        rows = {t: synth.table_rows(self.cat, t, self.sf) for t in self.cat.tables}
        return run_oracle(text, host_columns(self.cat, tpch.plan_columns(text), rows, sf=self.sf)), {"timeInMicrosecondsForPlan": 12.7}


def start(backend):
    srv = HTTPServer(("127.0.0.1", 0), serve.make_handler(backend))
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    return srv, f"http://127.0.0.1:{srv.server_address[1]}"


def post(url, body, ctype="text/vdl"):
    req = urllib.request.Request(url, data=body.encode(), headers={"Content-Type": ctype}, method="POST")
    try:
        with urllib.request.urlopen(req, timeout=60) as r:
            return r.status, r.read().decode()
    except urllib.error.HTTPError as e:
        return e.code, e.read().decode()


def test_post_of_the_printed_program_returns_the_servers_json(catalog):
    srv, base = start(OracleBackend(catalog))
    try:
        text = plan_text("q05.vdl")
        want = OracleBackend(catalog).run(text)[0]
        for path in serve.RUN_PATHS:
            code, body = post(base + path, text.replace("\n", " ;; Metadata {x = 1}\n", 3))      # a --metadata suffix is tolerated
            assert code == 200
            doc = json.loads(body)
            assert list(doc) == ["results", "timings"] and doc["timings"] == {"timeInMicrosecondsForPlan": 12}
            got = {next(iter(v)): next(iter(v.values())) for v in doc["results"].values()}
            assert got == {"." + k: [int(x) for x in v] for k, v in want.items()}
            cols = resolve.resolve(doc, catalog)                                                    # what ./resolve.py prints from it
            assert [n for n, _ in cols] == [".n_name", ".revenue"]                                  # alias of the dictionary column, whole name otherwise
    finally:
        srv.shutdown()


def test_errors_are_http_errors_with_the_message(catalog):
    srv, base = start(OracleBackend(catalog))
    try:
        assert post(base + "/voodoo/b200/run", "1,Load,t.a\n", ctype="text/voodoo")[0] == 415
        assert post(base + "/somewhere/else", "1,Load,t.a\n")[0] == 404
        code, body = post(base + "/voodoo/b200/run", "1,Load,lineitem.l_quantity\n2,Semisort,Id 1\n")
        assert code == 400 and "Semisort" in json.loads(body)["error"]
    finally:
        srv.shutdown()
