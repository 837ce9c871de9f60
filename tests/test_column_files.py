"""The reference's on-disk column format: 16-byte object header (mmod 0xfd) + raw payload.
Fixtures under tests/golden/colfiles/ were written by the reference itself (`(set "path" vec)`, see make_colfiles.rfl
there).  CPU tests: our reader maps them; our writer produces byte-identical files; the compiled reference (where
present) reads back what we write.  GPU test: a mapped file is folded through the host layer."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import bindings as ob
from rayforce_b200 import ColumnFile, capi

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "colfiles")
EXPECT = {"x_i64": (ob.I64, np.arange(1000, dtype=np.int64) - 500), "f_f64": (ob.F64, np.arange(10, dtype=np.float64)),
          "k_i32": (ob.I32, np.arange(7, dtype=np.int32))}


@pytest.mark.parametrize("name", sorted(EXPECT))
def test_reader_maps_files_written_by_the_reference(name):
    t, want = EXPECT[name]
    f = ColumnFile(os.path.join(HERE, name))
    assert (f.type, f.len) == (t, want.shape[0]) and np.array_equal(f.array, want)
    f.close()


@pytest.mark.parametrize("name", sorted(EXPECT))
def test_writer_is_byte_identical_to_the_reference(name, tmp_path):
    t, arr = EXPECT[name]
    p = str(tmp_path / name)
    ref = ColumnFile(os.path.join(HERE, name))
    ColumnFile.write(p, t, arr, attrs=ref.attrs)      # the reference stamps ATTR_ASC|ATTR_DISTINCT on `til`-derived columns
    ref.close()
    assert open(p, "rb").read() == open(os.path.join(HERE, name), "rb").read()


def test_reader_rejects_what_is_not_a_simple_column(tmp_path):
    p = str(tmp_path / "junk")
    open(p, "wb").write(b"\x00" * 64)
    with pytest.raises(capi.RfbError) as e:
        ColumnFile(p)
    assert e.value.kind == "type"
    open(p, "wb").write(b"\xfd\x00\x05\x00" + b"\x00" * 4 + (1000).to_bytes(8, "little") + b"\x00" * 8)   # claims 1000 rows, has 1
    with pytest.raises(capi.RfbError):
        ColumnFile(p)
    with pytest.raises(capi.RfbError):
        ColumnFile(str(tmp_path / "missing"))


def test_reference_reads_back_what_we_write(reference, tmp_path):
    p = str(tmp_path / "col")
    arr = (np.arange(50_000, dtype=np.int64) * 7919) % 1000 - 300
    ColumnFile.write(p, ob.I64, arr)
    r = reference.eval('(sum (get "%s"))' % p)
    assert int(reference.to_numpy(r)[0]) == int(arr.sum())


@pytest.mark.gpu
def test_fold_a_mapped_column_file_on_the_gpu(ctx, tmp_path):
    n = 3_000_017
    arr = (np.random.default_rng(1).integers(-1000, 1000, n)).astype(np.int64)
    arr[::101] = ob.NULL_I64
    p = str(tmp_path / "big")
    ColumnFile.write(p, ob.I64, arr)
    f = ColumnFile(p)
    got, nbytes = ctx.filter_fold_host(capi.GE, f.type, f.array, 0, capi.F_ALL, f.type, f.array)
    sel = arr[arr >= 0]
    assert nbytes == 8 * n and (got.rows, got.sum, got.min, got.max) == (sel.shape[0], int(sel.sum()), int(sel.min()), int(sel.max()))
    f.close()
