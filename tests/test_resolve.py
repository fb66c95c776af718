"""Result wire format: the server JSON shape resolve.py reads (resolve.py:8-32) and its dictionary decoding (64-94)."""
import json

import numpy as np

from mplan2vdl_b200 import resolve


def test_json_shape_and_dictionary_decoding(catalog):
    outputs = {"n_name__nation__n_name": np.array([72, 96, 1234], dtype=np.int64), "revenue": np.array([5, 6, 7], dtype=np.int64)}
    doc = json.loads(resolve.to_server_json(outputs, {"timeInMicrosecondsForFragment0": 215.7}))
    assert list(doc) == ["results", "timings"]
    assert doc["results"]["tmp0"] == {".n_name__nation__n_name": [72, 96, 1234]}
    assert doc["results"]["tmp1"] == {".revenue": [5, 6, 7]}
    assert doc["timings"] == {"timeInMicrosecondsForFragment0": 215}
    cols = resolve.resolve(doc, catalog)
    assert cols[0] == (".n_name", ["BRAZIL", "CANADA", 1234])        # dictionary.csv:3-4; unknown codes stay numeric
    assert cols[1] == (".revenue", [5, 6, 7])
    assert resolve.to_csv(cols) == ".n_name,.revenue\r\nBRAZIL,5\r\nCANADA,6\r\n1234,7\r\n"


def test_columns_of_different_length_are_padded(catalog):
    cols = resolve.resolve({"results": {"tmp0": {".a": [1, 2]}, "tmp1": {".l_returnflag__lineitem__l_returnflag": [16]}}}, catalog)
    assert resolve.to_csv(cols).splitlines()[2].endswith("-")


def test_decoder_matches_the_reference_byte_for_byte(catalog):
    """tests/golden/resolve_*.json -> .csv were produced by the reference's own resolve.py
    (tools/make_golden_resolve.py); the restated decoder must print the same bytes."""
    import glob
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    cases = sorted(glob.glob(os.path.join(here, "golden", "resolve_*.json")))
    assert len(cases) >= 2
    for path in cases:
        want = open(path[:-5] + ".csv", "rb").read()
        got = resolve.to_csv(resolve.resolve(open(path).read(), catalog)).encode()
        assert got == want, path
