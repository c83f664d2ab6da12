"""Known-answer tests that pin the oracle where the reference has no tests of its own:
analytic tone responses, Hann coherent gain, Parseval, quantiser values, derived sizes, index math,
and the oracle's stand-in FFT against numpy's float64 pocketfft."""
import numpy as np
import pytest

import oracle
from phantomsdr_b200 import sizes
from phantomsdr_b200 import SpectrumConfig

f32 = np.float32


@pytest.mark.parametrize("n", [4, 24, 360, 492, 548, 1024, 2520, 4096])
@pytest.mark.parametrize("sign", [-1, 1])
def test_generic_dft_matches_numpy(n, sign):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    want = np.fft.fft(x) if sign < 0 else np.fft.ifft(x) * n
    got64 = oracle.dft(x, sign)
    got32 = oracle.dft(x.astype(np.complex64), sign)
    assert np.abs(got64 - want).max() <= 1e-12 * np.abs(want).max()
    assert np.abs(got32 - want).max() <= 2e-6 * np.abs(want).max()


def test_forward_fft_vs_numpy_and_shadow():
    N = 1 << 14
    rng = np.random.default_rng(0)
    f = oracle.OracleFFT(N, oracle.downsample_levels(N, 1024), 0)
    f.set_output_additional_size(64)
    f.plan_c2c()
    a1 = (rng.standard_normal(N) * 1e-3).astype(f32)
    a2 = (rng.standard_normal(N) * 1e-3).astype(f32)
    f.load_complex_input(a1, a2)
    x = np.concatenate([a1.view(np.complex64), a2.view(np.complex64)]) * f.window
    assert np.array_equal(f.inbuf.view(np.complex64), x.astype(np.complex64))
    f.execute()
    want = np.fft.fft(x.astype(np.complex128)) / N
    assert np.abs(f.spectrum[:N] - want).max() <= 2e-6 * np.abs(want).max()
    assert np.abs(f.shadow_f64()[:N] - want).max() <= 1e-12 * np.abs(want).max()


@pytest.mark.parametrize("log2n,is_real", [(21, False), (22, True)])
def test_large_transforms_vs_numpy(log2n, is_real):
    """The GPU suite checks 2^21..2^23-point transforms against this oracle (tests/test_gpu_forward.py::
    test_large_transform_sizes): pin the oracle itself at those lengths, deep pyramid included (12+ levels)."""
    N = 1 << log2n
    R = N // 2 if is_real else N
    rng = np.random.default_rng(log2n)
    L = oracle.downsample_levels(R, 1024)
    f = oracle.OracleFFT(N, L, 0)
    f.set_output_additional_size(0 if is_real else 720)
    nfl = N // 2 if is_real else N
    a1 = (rng.standard_normal(nfl) * 1e-3).astype(f32)
    a2 = (rng.standard_normal(nfl) * 1e-3).astype(f32)
    if is_real:
        f.plan_r2c()
        f.load_real_input(a1, a2)
        x = np.concatenate([a1, a2]).astype(np.float64) * f.window
        want = np.fft.rfft(x)[:R] / N
    else:
        f.plan_c2c()
        f.load_complex_input(a1, a2)
        x = np.concatenate([a1.view(np.complex64), a2.view(np.complex64)]).astype(np.complex128) * f.window
        want = np.fft.fft(x) / N
    f.execute()
    assert np.abs(f.spectrum[:R] - want).max() <= 3e-6 * np.abs(want).max()
    q = f.quantized
    assert q.size == sum(R >> i for i in range(L)) and L >= 12
    # the last level is the pairwise-sum tree all the way down: its power sums are the block sums of |X|^2
    top = R >> (L - 1)
    disp = np.roll(np.abs(f.spectrum[:R]) ** 2, -(R // 2 + 1)) if not is_real else np.abs(f.spectrum[:R]) ** 2
    blocks = disp.reshape(top, -1).sum(axis=1)
    expect = 127 + 20 * np.log10(blocks) + 6.020599913 * (log2n - (L - 1))
    got = q[-top:].astype(np.int64)
    assert np.abs(got - expect).max() <= 1.5  # polynomial log2 (0.05 dB) + truncation


def test_r2c_vs_numpy_and_nyquist_left_unnormalised():
    N = 1 << 13
    rng = np.random.default_rng(1)
    f = oracle.OracleFFT(N, 3, 0)
    f.plan_r2c()
    a1 = (rng.standard_normal(N // 2) * 1e-3).astype(f32)
    a2 = (rng.standard_normal(N // 2) * 1e-3).astype(f32)
    f.load_real_input(a1, a2)
    x = np.concatenate([a1, a2]) * f.window
    f.execute()
    want = np.fft.rfft(x.astype(np.float64))
    got = f.spectrum
    assert np.abs(got[:N // 2] - want[:N // 2] / N).max() <= 2e-6 * np.abs(want / N).max()
    # src/fft_impl.cpp:152-154 divides only outbuf_len = N/2 bins: bin N/2 keeps the raw FFTW value
    assert abs(got[N // 2] - want[N // 2]) <= 2e-6 * np.abs(want).max()


def test_tone_known_answers_iq():
    """tone at FFT bin k, amplitude A: X[k]/N = A/2, X[k+-1]/N = -A/4, display index (k - N/2 - 1) mod N,
    q = trunc(127 + 20 log10((A/2)^2) + 6.0206 log2 N) within the polynomial's 0.05 dB."""
    N, k, A = 1 << 12, 777, 1e-2
    t = np.arange(N)
    x = (A * np.exp(2j * np.pi * k * t / N)).astype(np.complex64)
    L = oracle.downsample_levels(N, 1024)
    f = oracle.OracleFFT(N, L, 0)
    f.set_output_additional_size(16)
    f.plan_c2c()
    f.load_complex_input(x[:N // 2].view(f32), x[N // 2:].view(f32))
    f.execute()
    X = f.spectrum[:N]
    assert int(np.argmax(np.abs(X))) == k
    assert abs(abs(X[k]) - A / 2) < 1e-5 * A
    assert abs(X[k + 1] + A / 4) < 1e-5 * A and abs(X[k - 1] + A / 4) < 1e-5 * A
    q = f.quantized
    d = (k - N // 2 - 1) % N
    assert int(np.argmax(q[:N])) == d
    expect = 127 + 20 * np.log10((A / 2) ** 2) + 6.020599913 * np.log2(N)
    assert abs(int(q[d]) - expect) <= 1.0
    # level 1 holds the pairwise SUM (not max) with offset - 1: +3.01 dB - 6.02 dB vs the peak bin alone
    p = f.powerbuf
    assert np.array_equal(p[N:N + N // 2], p[0:N:2] + p[1:N:2])
    assert int(q[N + d // 2]) in (int(q[d]) - 5, int(q[d]) - 4, int(q[d]) - 3, int(q[d]) - 6)


def test_hann_properties_and_parseval():
    N = 1 << 12
    w = oracle.hann_window(N).astype(np.float64)
    assert abs(w.mean() - 0.5) < 1e-6  # coherent gain 0.5
    assert w[0] == 0.0 and abs(w[N // 2] - 1.0) < 1e-7  # periodic Hann
    rng = np.random.default_rng(2)
    f = oracle.OracleFFT(N, 1, 0)
    f.plan_c2c()
    a = (rng.standard_normal(2 * N)).astype(f32)
    f.load_complex_input(a[:N], a[N:])
    xin = f.inbuf.view(np.complex64).astype(np.complex128)
    f.execute()
    lhs = (np.abs(f.spectrum[:N].astype(np.complex128)) ** 2).sum()
    rhs = (np.abs(xin) ** 2).sum() / N
    assert abs(lhs - rhs) <= 1e-5 * rhs


def test_quantiser_values_and_wrap():
    q = oracle.lib().orc_quantize_one
    assert q(0.0, 20) == -128                      # log of zero clamps (std::max(-128.f, .))
    # NaN never reaches std::max as a NaN: vec_log2 only reads the bit pattern (exponent 255, mantissa 1.5)
    # -> (147 + 1.586) * 6.0206 + 127 = 1021 -> low byte 253 -> -3
    assert q(float("nan"), 20) == -3
    # p = 2^-20 at offset 20: log2 term = (127-20-128+20) + poly(1.0) = -1 + 1.00494 -> q = trunc(127.03)
    assert q(2.0 ** -20, 20) == 127
    # p = 2^-18 -> 139.07 dB: above int8: wraps modulo 256 like the x86 store (documented UB in the reference)
    assert q(2.0 ** -18, 20) == np.int8(np.uint8(139))
    assert q(2.0 ** -40, 20) == 6                  # trunc(127 + 6.0206 * (-20 - 1 + 1.00494))


def test_derived_sizes_match_reference_formulas():
    # configs 1-3
    assert oracle.audio_fft_size(12000, 1 << 17, 2_880_000) == 548
    assert oracle.audio_fft_size(12000, 1 << 17, 3_200_000) == 492
    assert oracle.audio_fft_size(12000, 1 << 20, 35_000_000) == 360
    assert oracle.audio_fft_size(12000, 1 << 21, 70_000_000) == 360
    assert oracle.downsample_levels(1 << 17) == 8 and oracle.downsample_levels(1 << 20) == 11
    assert oracle.skip_num(35_000_000, 1 << 20) == 6 and oracle.skip_num(2_880_000, 1 << 17) == 4
    assert oracle.skip_num(100_000, 1 << 17) == 1
    # the product's host logic restates the same formulas
    for sps, n, real in [(2_880_000, 1 << 17, False), (35_000_000, 1 << 20, False), (70_000_000, 1 << 21, True),
                         (3_200_000, 1 << 17, False), (1_000_000, 1 << 16, True)]:
        cfg = SpectrumConfig(sps=sps, fft_size=n, is_real=real)
        assert cfg.audio_fft_size == oracle.audio_fft_size(12000, n, sps)
        assert cfg.downsample_levels == oracle.downsample_levels(cfg.fft_result_size)
        assert cfg.skip_num == oracle.skip_num(sps, n)
        for l in (0, 1, n // 4, cfg.fft_result_size - 1):
            assert cfg.slice_offset(l) == oracle.slice_offset(l, n, real)
        for lv in range(cfg.downsample_levels):
            assert sizes.level_offset(lv, cfg.fft_result_size) == oracle.level_offset(lv, cfg.fft_result_size)


def test_sample_conversion():
    raw = np.array([0, 1, 127, 128, 129, 255], np.uint8)
    assert np.array_equal(oracle.convert(raw), np.array([-1, -127 / 128, -1 / 128, 0, 1 / 128, 127 / 128], f32))
    raw16 = np.array([0, 32768, 65535], np.uint16)
    assert np.array_equal(oracle.convert(raw16), np.array([-1, 0, 32767 / 32768], f32))
    assert np.array_equal(oracle.convert(np.array([-128, 0, 127], np.int8)), np.array([-1, 0, 127 / 128], f32))


def test_usb_tone_audio_frequency_and_ola_continuity():
    """a carrier d bins above floor(mid) demodulates (USB) to a tone of d/n cycles per output
    sample, continuous across odd/even frames for both parities of floor(mid) (the parity flip of
    signal.cpp:160-168)."""
    cfg = SpectrumConfig(sps=4_370_000, fft_size=1 << 14, audio_sps=12000 * 8)  # n = 360 at a small FFT
    N, n = cfg.fft_size, cfg.audio_fft_size
    assert n == 360
    for mid_bin in (3000, 3001):
        d = 20
        kdisp = mid_bin + d
        kfft = (kdisp + N // 2 + 1) % N
        nfr = 6
        t = np.arange((nfr + 1) * N // 2)
        x = (1e-2 * np.exp(2j * np.pi * kfft * t / N)).astype(np.complex64)
        f = oracle.OracleFFT(N, cfg.downsample_levels, 0)
        f.set_output_additional_size(n)
        f.plan_c2c()
        c = oracle.OracleClient(False, n, cfg.audio_sps, cfg.fft_result_size)
        assert c.on_window_message(mid_bin, float(mid_bin), mid_bin + 90)
        c.set_audio_demodulation(oracle.USB)
        audio = []
        for fr in range(nfr):
            a1 = x[fr * N // 2:(fr + 1) * N // 2]
            a2 = x[(fr + 1) * N // 2:(fr + 2) * N // 2]
            f.load_complex_input(a1.view(f32), a2.view(f32))
            f.execute()
            f.wrap_copy(n)
            ok, pcm, pwr, pre = c.send_audio(f.spectrum, N, fr)
            assert ok
            audio.append(pre)
        y = np.concatenate(audio[1:]).astype(np.float64)  # first frame only has half the overlap
        spec = np.abs(np.fft.rfft(y * np.hanning(y.size)))
        peak = np.argmax(spec) / y.size
        assert abs(peak - d / n) < 1.5 / y.size, (mid_bin, peak, d / n)
        # continuity: a clean tone has no energy far from its line (a sign error at frame joins would splatter)
        assert spec[np.argmax(spec)] > 200 * np.median(spec)


def test_window_validation_mirrors_reference():
    c = oracle.OracleClient(False, 360, 12000, 1 << 17)
    R = 1 << 17
    assert c.on_window_message(10, 20.0, 100)
    assert not c.on_window_message(-1, 0.0, 10)
    assert not c.on_window_message(10, 0.0, R)        # r >= R rejected (signal.cpp:304-306)
    assert not c.on_window_message(100, 0.0, 10)      # l > r
    assert not c.on_window_message(0, 0.0, 361)       # wider than audio_fft_size
    assert c.on_window_message(0, 0.0, 360)
    assert (c.l, c.r) == (0, 360)
