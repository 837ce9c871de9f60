"""The C-ABI boundary (CPU only, no compute): the library loads, exports every function include/rfb200.h declares, the
ctypes table covers them all, and the compute entry points fail loudly without a GPU instead of falling back."""
import ctypes as C
import os
import re

import pytest

from rayforce_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header="rfb200.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(rfb_[a-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if not n.endswith("_t")))


def test_library_is_built_and_exports_every_declared_symbol():
    assert os.path.exists(capi.LIB_PATH), "run __graft_entry__.build() first"
    lib = C.CDLL(capi.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 35
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, "declared in include/rfb200.h but not exported: %r" % missing


def test_ctypes_table_matches_header():
    assert sorted(capi.SIGNATURES) == declared_functions()


def test_abi_version_and_type_tables():
    lib = capi.load()
    assert lib.rfb_abi_version() == 1
    # result typing answers without a device (reference core/math.c:92-223 infer_*_type)
    assert lib.rfb_binop_type(capi.ADD, capi.I32, capi.I64) == capi.I64
    assert lib.rfb_binop_type(capi.DIV, capi.I32, capi.F64) == capi.I32      # `/` keeps the left operand's type
    assert lib.rfb_binop_type(capi.FDIV, capi.I64, capi.I64) == capi.F64
    assert lib.rfb_binop_type(capi.MOD, capi.I64, capi.I32) == capi.I32      # `%` takes the right operand's type
    assert lib.rfb_binop_type(capi.ADD, capi.U8, capi.I64) == capi.ERR_TYPE  # the plain-numeric typing rule; the full matrix is per form:
    assert lib.rfb_binop_type_form(capi.ADD, 0, capi.U8, capi.I64) == capi.I64
    assert lib.rfb_binop_type_form(capi.ADD, 0, capi.DATE, capi.TIME) == capi.TIMESTAMP   # core/math.c: date + time -> timestamp
    assert lib.rfb_binop_type_form(capi.SUB, 1, capi.TIMESTAMP, capi.TIMESTAMP) == capi.I64
    assert lib.rfb_binop_type_form(capi.MUL, 0, capi.DATE, capi.DATE) == capi.ERR_TYPE
    assert lib.rfb_binop_type_form(capi.ADD, 2, capi.I32, capi.I64) == lib.rfb_binop_type(capi.ADD, capi.I32, capi.I64)
    assert lib.rfb_aggr_type(capi.A_AVG, capi.I64) == capi.F64
    assert lib.rfb_aggr_type(capi.A_MIN, capi.I32) == capi.ERR_TYPE           # grouped min/max has no I32 case
    assert lib.rfb_aggr_type(capi.A_SUM, capi.TIMESTAMP) == capi.ERR_TYPE


def test_no_cpu_fallback_without_a_device():
    lib = capi.load()
    if lib.rfb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    rc = lib.rfb_ctx_create(0, C.byref(h))
    assert rc == capi.ERR_CUDA and not h.value
    assert b"no CPU fallback" in lib.rfb_last_error()
    from rayforce_b200 import Context, RfbError
    with pytest.raises(RfbError):
        Context(0)


def test_operator_layer_exports_every_declared_symbol():
    """include/rfb200_ops.h: the reference-facing operator layer (pure C) — same check, no compute"""
    from rayforce_b200 import ops
    assert os.path.exists(ops.OPS_PATH), "run __graft_entry__.build() first"
    capi.load()
    lib = C.CDLL(ops.OPS_PATH)
    names = [n for n in declared_functions("rfb200_ops.h") if n.startswith("rfb_") and not n.endswith("_p")]
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # every operator the Python harness binds is declared in the header
    for n in ops.UNARY + ops.BINARY:
        assert "rfb_" + n in names, n
    # the builtin host hands out objects with the reference's 16-byte header layout (core/rayforce.h:112-133)
    lib.rfb_ops_builtin_host.restype = C.POINTER(ops.HostApi)
    host = lib.rfb_ops_builtin_host().contents
    v = host.vector(capi.I64, 3)
    assert C.c_int8.from_address(v + 2).value == capi.I64 and C.c_uint32.from_address(v + 4).value == 1
    assert C.c_int64.from_address(v + 8).value == 3
    host.drop_obj(v)
    a = host.atom(capi.F64)
    assert C.c_int8.from_address(a + 2).value == -capi.F64
    host.drop_obj(a)
