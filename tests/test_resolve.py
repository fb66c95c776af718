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
    assert cols[0] == ("n_name", ["BRAZIL", "CANADA", 1234])        # dictionary.csv:3-4; unknown codes stay numeric
    assert cols[1] == ("revenue", [5, 6, 7])
    assert resolve.to_csv(cols).splitlines() == ["n_name,revenue", "BRAZIL,5", "CANADA,6", "1234,7"]


def test_columns_of_different_length_are_padded(catalog):
    cols = resolve.resolve({"results": {"tmp0": {".a": [1, 2]}, "tmp1": {".l_returnflag__lineitem__l_returnflag": [16]}}}, catalog)
    assert resolve.to_csv(cols).splitlines()[2].endswith("-")
