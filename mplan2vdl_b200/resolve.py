"""The result side of the executor boundary: the JSON the original Voodoo server returned and the decoding
`resolve.py` applied to it.

The reference pipeline (eval_query.sh:18-26) POSTed the program to ``/voodoo/cpu/run`` and piped the answer

    {"results": {"tmp66": {".o_orderpriority__orders__o_orderpriority": [16, 40, ...]},
                 "tmp75": {".order_count": [311, 263, ...]}},
     "timings": {"timeInMicrosecondsForFragment12": 215, ...}}                       (resolve.py:8-32)

into ``./resolve.py dictionary.csv``, which splits every output name ``<alias>__<table>__<column>``
(Vdl.hs:278-292) and, when ``<table>.<column>`` has dictionary entries, replaces the codes by their strings
(resolve.py:64-94), then prints the columns as CSV padded with ``-`` (resolve.py:101-107).  `to_server_json`
produces that JSON from `Plan.run()`'s ``{name: int64 array}`` (so the reference's own resolve.py can consume it
unchanged) and `resolve` / `to_csv` restate the decoder for Python 3.
"""
from __future__ import annotations

import csv
import io
import json

from .meta import Catalog


def to_server_json(outputs: dict, timings_us: dict | None = None) -> str:
    """`outputs` in MaterializeCompact order -> the server's JSON document.  One ``tmpN`` object per output, its single
    key the output name prefixed with '.', exactly what resolve.py:52-62 expects."""
    results = {f"tmp{i}": {"." + name: [int(v) for v in vals]} for i, (name, vals) in enumerate(outputs.items())}
    doc = {"results": results, "timings": {k: int(v) for k, v in (timings_us or {}).items()}}
    return json.dumps(doc)


def resolve(doc: dict | str, cat: Catalog) -> list:
    """[(column name, values)] with dictionary codes replaced by strings -- resolve.py:52-94, same rules and the same
    bytes: the key's leading '.' STAYS in every header (resolve.py:64,76: ``outputname = names[0]`` of
    ``".alias__table__column".split('__')``), names without exactly a ``__table__column`` origin or without a dictionary
    pass through whole, unknown codes stay numeric.  Pinned against the reference's own decoder by
    tests/golden/resolve_*.{json,csv} (tools/make_golden_resolve.py)."""
    if isinstance(doc, str):
        doc = json.loads(doc)
    cols = []
    for res in doc.get("results", {}).values():
        (k, vals), = res.items()
        vals = vals or []
        names = k.split("__")
        if len(names) != 3:
            cols.append((k, list(vals)))
            continue
        alias, table, column = names
        decoder = {code: s for s, code in cat.dictionary.get(f"{table}.{column}", {}).items()}
        if not decoder:
            cols.append((k, list(vals)))
            continue
        cols.append((alias, [decoder.get(v, v) for v in vals]))
    return cols


def to_csv(cols: list) -> str:
    """resolve.py:97-107: header row, then the rows, short columns padded with '-'; csv.writer's default dialect
    (``\r\n`` line ends), as the reference."""
    n = max((len(v) for _, v in cols), default=0)
    buf = io.StringIO()
    w = csv.writer(buf)
    w.writerow([name for name, _ in cols])
    for i in range(n):
        w.writerow([v[i] if i < len(v) else "-" for _, v in cols])
    return buf.getvalue()
