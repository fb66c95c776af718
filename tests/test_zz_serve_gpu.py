"""The HTTP handler of mplan2vdl_b200/serve.py with the real backend: programs POSTed as text/vdl run through libvdl_cuda and
the JSON matches the CPU oracle on the same synthetic tables."""
import json

import pytest

from mplan2vdl_b200 import serve
from test_serve import OracleBackend, post, start
from util import plan_text


@pytest.mark.gpu
def test_gpu_backend_behind_the_handler(catalog):
    backend = serve.GpuBackend(0.01)
    srv, base = start(backend)
    try:
        for q in ("q06.vdl", "q12.vdl", "q06.vdl"):            # the second Q6: its columns are resident already
            text = plan_text(q)
            code, body = post(base + "/voodoo/b200/run", text)
            assert code == 200, body
            got = {next(iter(v))[1:]: next(iter(v.values())) for v in json.loads(body)["results"].values()}
            want = OracleBackend(catalog).run(text)[0]
            assert got == {k: [int(x) for x in v] for k, v in want.items()}
        code, body = post(base + "/voodoo/b200/run", "1,Load,lineitem.l_quantity\n2,Semisort,Id 1\n")
        assert code == 400 and "Semisort" in json.loads(body)["error"]
    finally:
        srv.shutdown()
        backend.ctx.close()
