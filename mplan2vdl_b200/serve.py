"""`python -m mplan2vdl_b200.serve [--port 25472] [--sf 1]` -- the HTTP face of the executor, for the reference pipeline as it
stands (eval_query.sh:18-26): mplan2vdl's stdout is POSTed with `Content-Type: text/vdl` and the answer is the server's JSON
(resolve.py:8-32), which `./resolve.py dictionary.csv` decodes.

    ./tpchrun META plan.mplan | sed 's/;;.*//g' \\
      | curl -H "Content-Type: text/vdl" --data-binary @- http://localhost:25472/voodoo/b200/run | ./resolve.py META/dictionary.csv

Only `text/vdl` -- the program text mplan2vdl itself prints (Vdl.hs:410-477) -- is understood: the original pipeline converts it
to another server's `text/voodoo` and rewrites that in between (eval_query.sh:21-23); those two stages are not in the
reference repository, and the library's own planner does their work.  Errors: 400 with the library's
message for a malformed or unsupported program, 415 for another content type, 404 for another path.  The tables are the synthetic
TPC-H-shaped ones of `--sf`, generated on first use and kept resident; one request at a time (a context is single-threaded)."""
from __future__ import annotations

import argparse
import json
import re
import sys
import threading
import time
from http.server import BaseHTTPRequestHandler, HTTPServer

RUN_PATHS = ("/voodoo/b200/run", "/voodoo/cpu/run")      # the second: so the reference script's URL works unchanged


class GpuBackend:
    """Executes programs through libvdl_cuda on one GPU; columns are generated on first use and stay resident."""

    def __init__(self, sf: float, device: int = 0, fuse: bool = True):
        from . import tpch
        from .executor import Context
        from .meta import builtin_catalog
        self.sf, self.fuse, self.tpch = sf, fuse, tpch
        self.cat = builtin_catalog()
        self.ctx = Context(device)
        self.loaded = set()

    def run(self, text: str):
        cols = [c for c in self.tpch.plan_columns(text) if c not in self.loaded]
        self.tpch.load_synthetic(self.ctx, self.cat, cols, self.sf)
        self.loaded.update(cols)
        plan = self.ctx.plan(text, fuse=self.fuse)
        try:
            t0 = time.perf_counter()
            out = plan.run()
            return out, {"timeInMicrosecondsForPlan": 1e6 * (time.perf_counter() - t0)}
        finally:
            plan.close()


def make_handler(backend, lock=None):
    from . import resolve
    lock = lock or threading.Lock()

    class Handler(BaseHTTPRequestHandler):
        def _send(self, code: int, body: str, ctype: str = "application/json"):
            data = body.encode()
            self.send_response(code)
            self.send_header("Content-Type", ctype)
            self.send_header("Content-Length", str(len(data)))
            self.end_headers()
            self.wfile.write(data)

        def do_POST(self):          # noqa: N802 (http.server's naming)
            if self.path.rstrip("/") not in RUN_PATHS:
                return self._send(404, json.dumps({"error": f"POST {' or '.join(RUN_PATHS)}"}))
            ctype = (self.headers.get("Content-Type") or "").split(";")[0].strip().lower()
            if ctype != "text/vdl":
                return self._send(415, json.dumps({"error": "Content-Type must be text/vdl (the program mplan2vdl prints)"}))
            text = self.rfile.read(int(self.headers.get("Content-Length") or 0)).decode()
            text = re.sub(r" ;;.*", "", text)            # --metadata suffix (the pipeline strips it with sed; tolerated here)
            try:
                with lock:
                    out, timings = backend.run(text)
            except Exception as e:                       # the library's message: malformed / unsupported program, missing column
                return self._send(400, json.dumps({"error": str(e)}))
            self._send(200, resolve.to_server_json(out, timings) + "\n")

        def log_message(self, fmt, *args):
            sys.stderr.write("[vdl serve] " + fmt % args + "\n")

    return Handler


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m mplan2vdl_b200.serve")
    ap.add_argument("--port", type=int, default=25472, help="the original server's port (eval_query.sh:21)")
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--sf", type=float, default=1.0, help="scale factor of the synthetic tables")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--no-fuse", action="store_true", help="op-at-a-time execution only")
    args = ap.parse_args(argv)
    backend = GpuBackend(args.sf, args.device, fuse=not args.no_fuse)       # fails loudly without a GPU / the library
    srv = HTTPServer((args.host, args.port), make_handler(backend))
    sys.stderr.write(f"[vdl serve] listening on http://{args.host}:{args.port}{RUN_PATHS[0]} (SF {args.sf:g})\n")
    try:
        srv.serve_forever()
    except KeyboardInterrupt:
        pass
    return 0


if __name__ == "__main__":
    sys.exit(main())
