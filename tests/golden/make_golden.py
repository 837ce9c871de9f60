#!/usr/bin/env python
"""Generates tests/golden/reference_kat.json — known-answer vectors for the hot path, taken from the reference's OWN tests.

Run in the authoring container only (needs /root/reference and oracle/_ref/librayforce_ref.so):

    python tests/golden/make_golden.py

For every `TEST_ASSERT_EQ("(<op> <args>...)", "<expected>")` in the reference's tests/lang.c and tests/sort.c whose
operator is on the hot path and whose operands are plain numeric atoms/vectors, it
  * evaluates each operand expression and the expected-value expression with the compiled reference (eval_str),
  * checks — like the reference's test macro does (tests/main.c:124-143) — that the reference's result for the whole
    left-hand expression equals the expected value,
  * records operands and expected value as typed arrays (floats as IEEE bit patterns).
`TEST_ASSERT_ER(expr, "type"|"length")` lines for those operators are recorded as expected errors.
The committed JSON is what tests/test_oracle_golden.py pins the oracle against; nothing at test time reads the
reference tree.
"""
from __future__ import annotations

import base64
import ctypes as C
import gzip
import zlib
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as ob  # noqa: E402

REF_TESTS = "/root/reference/tests"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kat.json.gz")

UNARY = {"sum": "sum", "avg": "avg", "min": "min", "max": "max", "count": "count", "where": "where", "round": "round",
         "floor": "floor", "ceil": "ceil", "iasc": "iasc", "idesc": "idesc", "asc": "asc", "desc": "desc"}
BINARY = {"==": "eq", "!=": "ne", "<": "lt", ">": "gt", "<=": "le", ">=": "ge", "+": "add", "-": "sub", "*": "mul",
          "/": "div", "div": "fdiv", "%": "mod", "xbar": "xbar"}
NUMERIC = {ob.B8, ob.U8, ob.I16, ob.I32, ob.I64, ob.DATE, ob.TIME, ob.TIMESTAMP, ob.F64}


def split_top(expr: str):
    """'(op a b)' -> ['op', 'a', 'b'] (paren/bracket/brace/quote aware); None if not a call form"""
    s = expr.strip()
    if not (s.startswith("(") and s.endswith(")")):
        return None
    s = s[1:-1]
    out, depth, cur, instr = [], 0, "", False
    for ch in s:
        if instr:
            cur += ch
            if ch == '"':
                instr = False
            continue
        if ch == '"':
            instr = True
            cur += ch
        elif ch in "([{":
            depth += 1
            cur += ch
        elif ch in ")]}":
            depth -= 1
            cur += ch
        elif ch.isspace() and depth == 0:
            if cur:
                out.append(cur)
                cur = ""
        else:
            cur += ch
    if cur:
        out.append(cur)
    return out if depth == 0 else None


def obj_to_rec(R, o):
    """reference object -> {"type": t, "atom": bool, "values": [...]} or None when not a plain numeric atom/vector"""
    t = R.type_of(o)
    if t == 127:
        return None
    at = -t if t < 0 else t
    if at not in NUMERIC:
        return None
    v, _ = R.to_numpy(o, drop=False)
    arr = np.atleast_1d(np.asarray(v))
    as_i64 = arr.view(np.int64) if at == ob.F64 else arr.astype(np.int64)
    if as_i64.shape[0] > 256:   # long vectors: wrapping first differences, zlib, base64 (see tests/golden_io.py)
        with np.errstate(over="ignore"):
            d = np.diff(as_i64, prepend=np.int64(0))
        return {"type": at, "atom": False, "n": int(as_i64.shape[0]),
                "delta_zlib_b64": base64.b64encode(zlib.compress(d.astype("<i8").tobytes(), 9)).decode()}
    return {"type": at, "atom": t < 0, "values": [int(x) for x in as_i64]}


def fmt(R, o) -> str:
    """the reference's own rendering (obj_fmt(obj, B8_TRUE), what tests/main.c:127 compares)"""
    R.L.obj_fmt.restype = C.c_void_p
    R.L.obj_fmt.argtypes = [C.c_void_p, C.c_int64]
    s = R.L.obj_fmt(o, 1)
    n = R.len_of(s)
    txt = C.string_at(s + 16, n).decode(errors="replace")
    R.drop(s)
    return txt


def main():
    R = ob.Reference.get()
    cases, skipped = [], 0
    pat_eq = re.compile(r'TEST_ASSERT_EQ\(\s*"((?:[^"\\]|\\.)*)"\s*,\s*"((?:[^"\\]|\\.)*)"\s*\)')
    pat_er = re.compile(r'TEST_ASSERT_ER\(\s*"((?:[^"\\]|\\.)*)"\s*,\s*"((?:[^"\\]|\\.)*)"\s*\)')
    for fname in ("lang.c", "sort.c"):
        with open(os.path.join(REF_TESTS, fname)) as f:
            lines = f.readlines()
        for ln, line in enumerate(lines, 1):
            if line.lstrip().startswith("//"):
                continue
            for m in pat_eq.finditer(line):
                lhs, rhs = (x.encode().decode("unicode_escape") for x in m.groups())
                parts = split_top(lhs)
                if not parts:
                    continue
                op = parts[0]
                if not ((op in UNARY and len(parts) == 2) or (op in BINARY and len(parts) == 3)):
                    continue
                args = []
                for a in parts[1:]:
                    o = R.eval(a)
                    rec = obj_to_rec(R, o)
                    R.drop(o)
                    args.append(rec)
                eo = R.eval(rhs)
                etxt = fmt(R, eo)
                R.drop(eo)
                lo = R.eval(lhs)
                actual = obj_to_rec(R, lo)
                atxt = fmt(R, lo)
                R.drop(lo)
                if any(a is None for a in args) or actual is None:
                    skipped += 1
                    continue
                if atxt != etxt:   # the reference's own pass criterion (string equality of the renderings)
                    print("reference fails its own golden?", fname, ln, lhs, rhs, atxt, etxt, file=sys.stderr)
                    skipped += 1
                    continue
                # `expect` = the reference's actual result, which its test certifies renders as `expected_text`
                cases.append({"src": "tests/%s:%d" % (fname, ln), "expr": lhs, "expected_text": rhs,
                              "op": UNARY.get(op) or BINARY[op], "args": args, "expect": actual})
            for m in pat_er.finditer(line):
                lhs, err = m.groups()
                parts = split_top(lhs)
                if not parts or err not in ("type", "length"):
                    continue
                op = parts[0]
                if not ((op in UNARY and len(parts) == 2) or (op in BINARY and len(parts) == 3)):
                    continue
                args = []
                for a in parts[1:]:
                    o = R.eval(a)
                    rec = obj_to_rec(R, o)
                    R.drop(o)
                    args.append(rec)
                if any(a is None for a in args):
                    skipped += 1
                    continue
                cases.append({"src": "tests/%s:%d" % (fname, ln), "expr": lhs, "op": UNARY.get(op) or BINARY[op],
                              "args": args, "error": err})
    with gzip.open(OUT, "wt") as f:
        json.dump({"reference_commit": "2151d51d", "generator": "tests/golden/make_golden.py",
                   "note": "F64 values are IEEE-754 bit patterns as signed 64-bit integers", "cases": cases}, f,
                  separators=(",", ":"))
    by = {}
    for c in cases:
        by[c["op"]] = by.get(c["op"], 0) + 1
    print("wrote %d cases (%d skipped: non-numeric operands) -> %s" % (len(cases), skipped, OUT))
    print(sorted(by.items()))


if __name__ == "__main__":
    main()
