"""reader for tests/golden/reference_kat.json.gz (written by tests/golden/make_golden.py)"""
from __future__ import annotations

import base64
import gzip
import json
import os
import zlib

import numpy as np

from oracle import bindings as ob

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kat.json.gz")


def decode(rec):
    """-> (type, is_atom, numpy array of the type's dtype)"""
    t = rec["type"]
    if "values" in rec:
        raw = np.array(rec["values"], dtype=np.int64)
    else:
        d = np.frombuffer(zlib.decompress(base64.b64decode(rec["delta_zlib_b64"])), dtype="<i8")
        with np.errstate(over="ignore"):
            raw = np.cumsum(d, dtype=np.int64)
        assert raw.shape[0] == rec["n"]
    arr = raw.view(np.float64) if t == ob.F64 else raw.astype(ob.NP_OF[t])
    return t, bool(rec["atom"]), arr


def load_cases():
    with gzip.open(PATH, "rt") as f:
        return json.load(f)["cases"]
