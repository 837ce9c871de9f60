#!/usr/bin/env python
"""Extract the TYPE MATRIX of the reference's element-wise arithmetic — which (operator, operand types, vector/atom form) exist
and, for each, the four type roles of its kernel macro — from the switch tables of core/math.c (ray_add_partial ..
ray_xbar_partial, core/math.c:251-1782), and write it as a C table:
    rayforce_b200/csrc/binop_matrix.inc     (the device layer's dispatch: rfb_binop_dev / rfb_binop_type_form)
    rayforce_b200/csrc/binop_kernels.inc    (the kernel instantiations that dispatch needs)
    oracle/binop_matrix.inc                 (the oracle's restatement: rfo_binop_form)
Every case of the reference has the shape  out[i] = mt_to_ot(OP(lt_to_mt(x[i]), rt_to_mt(y[i])))  (__BINOP_V_V / _V_A / _A_V,
core/math.c:43-90) with (lt, rt, ot, mt, OP) spelled out per case; the table records exactly those five facts per case — typing
facts of the reference's interface, like the golden vectors (no code).  Run in the authoring container (needs /root/reference);
the generated files are committed."""
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = [("ADD", "ray_add_partial"), ("SUB", "ray_sub_partial"), ("MUL", "ray_mul_partial"), ("DIV", "ray_div_partial"),
       ("FDIV", "ray_fdiv_partial"), ("MOD", "ray_mod_partial"), ("XBAR", "ray_xbar_partial")]
TYPE = {"B8": 1, "U8": 2, "I16": 3, "I32": 4, "I64": 5, "DATE": 7, "TIME": 8, "TIMESTAMP": 9, "F64": 10}
KIND = {"b8": 1, "u8": 2, "i16": 3, "i32": 4, "i64": 5, "date": 7, "time": 8, "timestamp": 9, "f64": 10}
FORM = {"V_V": 0, "V_A": 1, "A_V": 2}
CLASS = {"U8": 2, "I16": 3, "I32": 4, "I64": 5, "F64": 10}


SIZE = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 7: 4, 8: 4, 9: 8, 10: 8}


def infer_table(src, fn):
    """(x type, y type) -> the element type binop_map gives the result vector (infer_math_type & co, core/math.c:92-249)"""
    m = re.search(r"i8_t %s\(obj_p x, obj_p y\) \{(.*?)\n\}\n" % fn, src, re.S)
    table, labels, default = {}, [], None
    for line in m.group(1).split("\n"):
        line = line.strip()
        c = re.match(r"case MTYPE2\(TYPE_(\w+), TYPE_(\w+)\):", line)
        if c:
            labels.append((c.group(1), c.group(2)))
            continue
        r = re.match(r"return TYPE_(\w+);", line)
        if r:
            if labels:
                for a, b in labels:
                    if a in TYPE and b in TYPE:
                        table[(TYPE[a], TYPE[b])] = TYPE[r.group(1)]
                labels = []
            else:
                default = TYPE[r.group(1)]
    return lambda xt, yt: table.get((xt, yt), default)


def main():
    src = open(os.path.join(REF, "core", "math.c")).read()
    infer = {n: infer_table(src, n) for n in ("infer_math_type", "infer_div_type", "infer_mod_type", "infer_xbar_type")}
    # binop_map (core/math.c:2295-2299): fdiv -> F64, div -> infer_div_type, xbar -> infer_xbar_type, mod -> infer_mod_type, else math
    vec_type = [infer["infer_math_type"], infer["infer_math_type"], infer["infer_math_type"], infer["infer_div_type"],
                lambda xt, yt: TYPE["F64"], infer["infer_mod_type"], infer["infer_xbar_type"]]
    rows = []
    dropped = 0
    for opi, (op, fn) in enumerate(OPS):
        m = re.search(r"obj_p %s\(obj_p x, obj_p y, i64_t len, i64_t offset, obj_p out\) \{(.*?)\n\}\n" % fn, src, re.S)
        line0 = src[:m.start()].count("\n") + 1
        labels = []
        for k, line in enumerate(m.group(1).split("\n")):
            line = line.strip()
            c = re.match(r"case MTYPE2\((-?)TYPE_(\w+), (-?)TYPE_(\w+)\):", line)
            if c:
                labels.append((c.group(1) == "-", c.group(2), c.group(3) == "-", c.group(4)))
                continue
            r = re.match(r"return __BINOP_(V_V|V_A|A_V)\(x, y, (\w+), (\w+), (\w+), (\w+), (\w+), len, offset, out\);", line)
            if r:
                form, lt, rt, ot, mt, name = r.groups()
                fam = re.match(r"(ADD|SUB|MUL|DIV|FDIV|MOD|XBAR)(U8|I16|I32|I64|F64)$", name)
                assert fam and fam.group(1) == op, (op, name)
                for xa, xt, ya, yt in labels:
                    assert (xa, ya) == {"V_V": (False, False), "V_A": (False, True), "A_V": (True, False)}[form], (op, labels, form)
                    if xt in TYPE and yt in TYPE:
                        vt = vec_type[opi](TYPE[xt], TYPE[yt])
                        if SIZE[vt] != SIZE[KIND[ot]]:
                            dropped += 1        # the kernel writes ot-sized elements into a vt-sized vector: not a defined result
                            continue
                        rows.append((opi, FORM[form], TYPE[xt], TYPE[yt], KIND[lt], KIND[rt], KIND[ot], KIND[mt], CLASS[fam.group(2)], vt, line0 + k))
                labels = []
                continue
            if line.startswith("return") or line.startswith("default"):
                labels = []
    rows.sort()
    text = ["/* GENERATED by tools/gen_binop_matrix.py from the switch tables of the reference's core/math.c (ray_add_partial .. ray_xbar_partial).",
            " * One entry per case of the reference:  out[i] = mt_to_ot(OP(lt_to_mt(x[i]), rt_to_mt(y[i])))",
            " * { op (0 add 1 sub 2 mul 3 div 4 fdiv 5 mod 6 xbar), form (0 vec-vec 1 vec-atom 2 atom-vec), x type, y type,",
            " *   lt, rt (how the operand storage is read), ot (result element type), mt (type the operator works in),",
            " *   operator family width (2 U8, 3 I16, 4 I32, 5 I64, 10 F64: ADDI32, MULF64, ...),",
            " *   vt = element type binop_map gives the result vector (infer_math_type & co, core/math.c:92-249, :2295-2299), core/math.c line }",
            " * %d cases; %d more cases of the reference are left out because their kernel writes elements of another width than the" % (len(rows), dropped),
            " * vector binop_map allocated (no defined result to match). */"]
    for r in rows:
        text.append("{%d, %d, %2d, %2d, %2d, %2d, %2d, %2d, %2d, %2d, %4d}," % r)
    body = "\n".join(text) + "\n"
    for rel in (("rayforce_b200", "csrc", "binop_matrix.inc"), ("oracle", "binop_matrix.inc")):
        with open(os.path.join(ROOT, *rel), "w") as f:
            f.write(body)
    # the device layer instantiates one kernel per (operand storage, operand storage, result storage, operator family, operator)
    # that occurs outside I32/I64/F64 x I32/I64/F64 (those keep k_map.cu's kernels)
    store = {1: "u8", 2: "u8", 3: "i16", 4: "i32", 5: "i64", 7: "i32", 8: "i32", 9: "i64", 10: "f64"}
    plain = (TYPE["I32"], TYPE["I64"], TYPE["F64"])
    kernels = sorted({(store[r[4]], store[r[5]], store[r[6]], r[8], r[0]) for r in rows if not (r[2] in plain and r[3] in plain)})
    with open(os.path.join(ROOT, "rayforce_b200", "csrc", "binop_kernels.inc"), "w") as f:
        f.write("/* GENERATED by tools/gen_binop_matrix.py: TYPED(x storage, y storage, result storage, operator family, operator) for every\n"
                " * combination the reference's type matrix needs outside I32/I64/F64 operands — %d kernels */\n" % len(kernels))
        for k in kernels:
            f.write("TYPED(%s, %s, %s, %d, %d)\n" % k)
    print("%d cases, %d dropped, %d kernels" % (len(rows), dropped, len(kernels)))


if __name__ == "__main__":
    main()
