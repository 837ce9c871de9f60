"""Host-side model of the one-atomic accumulators of the narrow group-by's accumulate pass (k_fused_group.cu: sacc_add_packed /
sacc_flush_packed).  A slot's first 32-bit word packs the record count (top 8 bits) over the value sum (low 24 bits); a record
adds 2^24 + v with one returning atomic add, the adder that overflowed the sum field counts a carry, the adder that wrapped the
word counts a wrap.  Atomic adds on one word are totally ordered, so a sequential replay IS the device semantics: the flush
formulas must give the exact (sum, count) for any sequence of values below 2^20 (VB <= 20 bits, RecFmt<u32, KPL>)."""
import numpy as np
import pytest

M32 = (1 << 32) - 1


def replay(values):
    word = carries = wraps = 0
    for v in values:
        v = int(v)
        inc = (1 << 24) + v
        old = word
        word = (old + inc) & M32
        if (old & 0xFFFFFF) + v >= 1 << 24:
            carries += 1
        if ((old + inc) & M32) < inc:
            wraps += 1
    return word, carries, wraps


def flush(word, carries, wraps):
    total = (wraps << 32) | word
    return (carries << 24) + (word & 0xFFFFFF), (total >> 24) - carries


@pytest.mark.parametrize("case", ["random20", "max20", "zeros", "random19", "mixed", "short"])
def test_flush_recovers_exact_sum_and_count(case):
    r = np.random.default_rng(len(case))
    n = 200_000
    vals = {"random20": r.integers(0, 1 << 20, n), "max20": np.full(n, (1 << 20) - 1), "zeros": np.zeros(n, np.int64),
            "random19": r.integers(0, 1 << 19, n), "mixed": np.where(r.random(n) < 0.5, 0, (1 << 20) - 1),
            "short": r.integers(0, 1 << 20, 7)}[case]
    # every prefix length that matters: before / at / after the first carry and the first wrap, and the whole sequence
    for m in sorted({1, 2, 15, 16, 17, 31, 255, 256, 257, 4095, 65_537, len(vals)}):
        if m > len(vals):
            continue
        s, c = flush(*replay(vals[:m]))
        assert (s, c) == (int(vals[:m].sum()), m), (case, m)


def test_carry_and_wrap_are_rare():
    """the point of the layout: ~1/32 carries and ~1/250 wraps per record for uniform 20-bit values"""
    r = np.random.default_rng(1)
    n = 100_000
    _, carries, wraps = replay(r.integers(0, 1 << 20, n))
    assert carries < n / 24 and wraps < n / 200
