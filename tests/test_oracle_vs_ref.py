"""Pins the oracle's restatements bit-for-bit against the REFERENCE's own compiled sources
(oracle/_ref/libphantom_ref.so = src/utils/dsp.cpp + src/utils/audioprocessing.cpp + src/utils.h built
by oracle/Makefile from /root/reference). Skipped only if that library was never built."""
import numpy as np
import pytest

import oracle

ref = oracle.ref()
pytestmark = pytest.mark.skipif(ref is None, reason="oracle/_ref not built (needs /root/reference at build time)")
f32 = np.float32


def test_hann_window_bit_exact():
    for n in (4, 360, 1024, 1 << 16, 1 << 17):
        w = oracle.hann_window(n)
        w2 = np.empty(n, f32)
        ref.ref_build_hann_window(w2, n)
        assert np.array_equal(w, w2)


def test_fm_discriminator_bit_exact():
    rng = np.random.default_rng(1)
    for n in (1, 7, 180, 274):
        z = (rng.standard_normal(2 * n) * 10 ** rng.uniform(-4, 0)).astype(f32)
        a, b = np.empty(n, f32), np.empty(n, f32)
        oracle.lib().orc_polar_discriminator_fm(z, 0.3, -0.2, a, n)
        ref.ref_polar_discriminator_fm(z.copy(), 0.3, -0.2, b, n)
        assert np.array_equal(a, b)


def test_am_demod_and_int16_bit_exact():
    rng = np.random.default_rng(2)
    z = rng.standard_normal(2 * 300).astype(f32)
    a, b = np.empty(300, f32), np.empty(300, f32)
    oracle.lib().orc_am_demod(z, a, 300)
    ref.ref_dsp_am_demod(z.copy(), b, 300)
    assert np.array_equal(a, b)
    x = (rng.standard_normal(5000) * 1.5).astype(f32)
    x[:6] = [0.0, -0.0, 1.99996, -2.0, 2.5, -3.0]
    ia, ib = np.empty(5000, np.int32), np.empty(5000, np.int32)
    oracle.lib().orc_float_to_int16(x, ia, 16384.0, x.size)
    ref.ref_dsp_float_to_int16(x.copy(), ib, 16384.0, x.size)
    assert np.array_equal(ia, ib)
    assert ia.max() == 32767 and ia.min() == -32768


def test_negate_add_helpers_match_plain_ops():
    rng = np.random.default_rng(3)
    a = rng.standard_normal(360).astype(f32)
    b = rng.standard_normal(360).astype(f32)
    x = a.copy()
    ref.ref_dsp_negate_float(x, x.size)
    assert np.array_equal(x, -a)
    y = a.copy()
    ref.ref_dsp_add_float(y, b.copy(), 180)
    assert np.array_equal(y[:180], a[:180] + b[:180]) and np.array_equal(y[180:], a[180:])
    z = a.copy()
    ref.ref_dsp_add_complex(z, b.copy(), 90)
    assert np.array_equal(z[:180], a[:180] + b[:180])
    w = a.copy()
    ref.ref_dsp_negate_complex(w, 180)
    assert np.array_equal(w, -a)


@pytest.mark.parametrize("sr,frame", [(12000, 180), (12000, 274), (44100, 512), (8000, 13)])
def test_agc_bit_exact(sr, frame):
    rng = np.random.default_rng(4)
    a = oracle.OracleAGC(0.2, 50.0, 300.0, 200.0, float(sr))
    r = ref.ref_agc_create(0.2, 50.0, 300.0, 200.0, float(sr))
    for it in range(80):
        amp = 10 ** rng.uniform(-4, 0.3)
        x = (rng.standard_normal(frame) * amp).astype(f32)
        if it == 20:
            x[:] = 0  # silence: desired gain 0.2 / 1e-10
        if it == 40:
            a.reset()
            ref.ref_agc_reset(r)
        y = a.process(x)
        y2 = x.copy()
        ref.ref_agc_process(r, y2, y2.size)
        assert np.array_equal(y, y2), f"frame {it}"
    ref.ref_agc_destroy(r)


@pytest.mark.parametrize("delay", [32, 116, 2])
def test_dc_blocker_bit_exact(delay):
    rng = np.random.default_rng(5)
    d = oracle.OracleDC(delay)
    r = ref.ref_dc_create(delay)
    for it in range(40):
        x = (rng.standard_normal(180) * 0.3 + 0.25).astype(f32)
        y = d.remove(x)
        y2 = x.copy()
        ref.ref_dc_remove(r, y2, y2.size)
        assert np.array_equal(y, y2), f"frame {it}"
    ref.ref_dc_destroy(r)


def test_slice_power_is_re2_plus_im2():
    """std::norm on complex<float> is re*re + im*im in libstdc++ (not abs()^2); the oracle's pwr accumulation restates
    that (src/signal.cpp:117-119). Bit-exact."""
    rng = np.random.default_rng(6)
    z = (rng.standard_normal(2 * 90) * 1e-3).astype(f32)
    want = ref.ref_slice_power(z.copy(), 90)
    acc = f32(0)
    for i in range(90):
        acc = f32(acc + f32(f32(z[2 * i] * z[2 * i]) + f32(z[2 * i + 1] * z[2 * i + 1])))
    assert want == acc
