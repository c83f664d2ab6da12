"""CPU: the host-built tables of the table-driven waterfall quantiser (b200_quant_table, used by the opt-in
B200_OPT_PACKED_MATH bit 1 path) against the oracle's quantiser (vec_log2 + power_and_quantize,
src/fft_impl.cpp:14-44): every input outside a cell's [lo, hi) band must quantise to base / base + 1, the bands must
be tiny, and inside them the device takes the exact arithmetic anyway. tools/quant_table.c runs the same check over
ALL 2^31 inputs (offsets 0..48, about a minute); here: every cell edge, every band edge and a dense random sample."""
import ctypes as C

import numpy as np
import pytest

import oracle
from phantomsdr_b200.backend import quant_table


def _oracle_q(bits: np.ndarray, off: int) -> np.ndarray:
    L = oracle.lib()
    L.orc_quantize_one.restype = C.c_int8
    L.orc_quantize_one.argtypes = [C.c_float, C.c_int]
    vals = bits.astype(np.uint32).view(np.float32)
    return np.array([L.orc_quantize_one(float(v), off) & 0xFF for v in vals], dtype=np.int64)


def _table_q(bits: np.ndarray, lo, hi, base):
    c = (bits >> 20).astype(np.int64)
    inside = (bits >= lo[c]) & (bits < hi[c])
    q = (base[c].astype(np.int64) + (bits >= hi[c])) & 0xFF
    return q, inside


@pytest.mark.parametrize("off", [10, 17, 20, 21, 23, 31])
def test_table_reproduces_the_oracle_quantiser(off):
    lo, hi, base = quant_table(off)
    assert (hi >= lo).all() and int((hi - lo).max()) <= 8 and int((hi - lo).sum()) < 400
    cells = np.arange(2048, dtype=np.int64)
    edges = np.concatenate([cells << 20, ((cells + 1) << 20) - 1])                       # first / last value of every cell
    near = np.concatenate([lo.astype(np.int64) + d for d in range(-6, 7)] + [hi.astype(np.int64) + d for d in range(-6, 7)])
    rng = np.random.default_rng(off)
    rand = rng.integers(0, 1 << 31, size=60000, dtype=np.int64)
    # the interesting range: powers 1e-13 .. 1 (exponents ~84..127), where real spectra live
    live = rng.integers(84 << 23, 128 << 23, size=60000, dtype=np.int64)
    bits = np.unique(np.concatenate([edges, near, rand, live]))
    bits = bits[(bits >= 0) & (bits < (1 << 31))]
    # signalling-NaN payloads do not survive the double -> float conversion of the ctypes call (the hardware quiets
    # them), so they cannot be handed to the oracle; |X|^2 is never one
    bits = bits[(bits <= 0x7F800000) | (bits >= 0x7FC00000)].astype(np.uint32)
    want = _oracle_q(bits, off)
    got, inside = _table_q(bits, lo, hi, base)
    bad = (got != want) & ~inside
    assert not bad.any(), f"offset {off}: {int(bad.sum())} mismatches, first at bits {bits[bad][:4]}"
    assert inside.sum() <= (hi - lo).sum()


def test_sign_bit_reaches_the_polynomial_like_the_reference():
    # fft_impl.cpp:19 clears only the exponent field, so a (never occurring) negative power feeds a negative mantissa
    # to the polynomial: the tables cover sign = 0 only and the device sends sign = 1 down the exact path
    L = oracle.lib()
    L.orc_quantize_one.restype = C.c_int8
    L.orc_quantize_one.argtypes = [C.c_float, C.c_int]
    assert L.orc_quantize_one(1e-6, 20) != L.orc_quantize_one(-1e-6, 20)
