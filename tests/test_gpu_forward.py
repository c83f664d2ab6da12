"""GPU parity: forward FFT + waterfall pyramid through the C ABI vs the CPU oracle.

Tolerances: spectrum max|d| <= 1e-5 * max|X| per frame (float32 FFTs of different
factorisation agree to ~1e-6); the int8 pyramid is bit-exact when the oracle's quantiser is fed the
engine's own spectrum (integer/bit-trick work), and >= 99.9 % identical / |d| <= 1 against the
oracle's independent float32 FFT.
"""
import numpy as np
import pytest

import oracle
from phantomsdr_b200 import SpectrumConfig
from phantomsdr_b200.synth import SignalSource
from helpers import make_engine, make_oracle_fft, hop_as_floats, spectrum_tolerance_check

pytestmark = pytest.mark.gpu

CONFIGS = [
    pytest.param(SpectrumConfig(sps=2_880_000, fft_size=1 << 17, is_real=False), id="cfg1-iq-2^17"),
    pytest.param(SpectrumConfig(sps=2_000_000, fft_size=1 << 16, is_real=False), id="iq-2^16"),
    pytest.param(SpectrumConfig(sps=8_000_000, fft_size=1 << 18, is_real=False), id="iq-2^18"),
    pytest.param(SpectrumConfig(sps=16_000_000, fft_size=1 << 19, is_real=False), id="iq-2^19"),
    pytest.param(SpectrumConfig(sps=35_000_000, fft_size=1 << 20, is_real=False), id="cfg2-iq-2^20"),
    pytest.param(SpectrumConfig(sps=4_000_000, fft_size=1 << 17, is_real=True), id="real-2^17"),
    pytest.param(SpectrumConfig(sps=70_000_000, fft_size=1 << 21, is_real=True), id="cfg3-real-2^21"),
]


def _run_frames_against_oracle(cfg, next_hop, nframes):
    """Drive engine and oracle with the reference's ring order (src/fft.cpp:47-105) and compare every frame."""
    orc = make_oracle_fft(cfg)
    eng = make_engine(cfg)
    ring = [eng.malloc(cfg.hop_floats) for _ in range(3)]
    ring[0][:] = hop_as_floats(next_hop())
    ring[1][:] = hop_as_floats(next_hop())
    idx = 0
    R = cfg.fft_result_size
    for frame in range(nframes):
        a1, a2 = ring[idx], ring[(idx + 1) % 3]
        if cfg.is_real:
            eng.load_real_input(a1, a2)
            orc.load_real_input(a1, a2)
        else:
            eng.load_complex_input(a1, a2)
            orc.load_complex_input(a1, a2)
        ring[(idx + 2) % 3][:] = hop_as_floats(next_hop())  # the async read of fft.cpp:56-67
        idx = (idx + 1) % 3
        eng.execute()
        orc.execute()
        nb = R + 1 if cfg.is_real else R
        got = eng.get_output_buffer().view(np.complex64)
        ref = orc.spectrum
        spectrum_tolerance_check(got[:R], ref[:R])
        if cfg.is_real:
            # Nyquist bin stays unnormalised in the reference (src/fft_impl.cpp:152-154)
            assert abs(got[R] - ref[R]) <= 1e-5 * cfg.fft_size * np.abs(ref[:R]).max()
        else:
            n = cfg.audio_fft_size  # engine fills the wrap tail itself (src/fft.cpp:96-97)
            assert np.array_equal(got[R:R + n], got[:n])
        # f64 shadow arbitrates both float32 implementations
        shadow = orc.shadow_f64()
        spectrum_tolerance_check(got[:R], shadow[:R])
        # integer pipeline: bit-exact on the engine's own spectrum
        q_gpu = eng.get_quantized_buffer().copy()
        q_same = orc.requantize_from(got[:nb] if cfg.is_real else got[:R])
        assert q_gpu.shape == q_same.shape
        assert np.array_equal(q_gpu, q_same), f"pyramid differs at {np.flatnonzero(q_gpu != q_same)[:8]}"
        # and against the independent oracle FFT: identical except rounding-boundary crossings
        orc.load_real_input(a1, a2) if cfg.is_real else orc.load_complex_input(a1, a2)
        orc.execute()
        q_ref = orc.quantized
        d = np.abs(q_gpu.astype(np.int32) - q_ref.astype(np.int32))
        assert d.max() <= 1, f"pyramid off by {d.max()}"
        assert (d != 0).mean() <= 1e-3, f"{(d != 0).mean():.2e} of pyramid bytes differ"
    for b in ring:
        eng.free(b)
    eng.close()


@pytest.mark.parametrize("cfg", CONFIGS)
def test_forward_and_pyramid_match_oracle(gpu_required, cfg):
    src = SignalSource(cfg, seed=0x5EED + cfg.fft_size % 97)
    _run_frames_against_oracle(cfg, src.next_hop, 3 if cfg.fft_size <= (1 << 19) else 2)


LARGE = [
    pytest.param(SpectrumConfig(sps=35_000_000, fft_size=1 << 21, is_real=False), id="iq-2^21"),
    pytest.param(SpectrumConfig(sps=35_000_000, fft_size=1 << 22, is_real=False), id="iq-2^22"),
    pytest.param(SpectrumConfig(sps=35_000_000, fft_size=1 << 23, is_real=False), id="iq-2^23"),
    pytest.param(SpectrumConfig(sps=70_000_000, fft_size=1 << 22, is_real=True), id="real-2^22"),
    pytest.param(SpectrumConfig(sps=70_000_000, fft_size=1 << 23, is_real=True), id="real-2^23"),
]


@pytest.mark.parametrize("cfg", LARGE)
def test_large_transform_sizes(gpu_required, cfg):
    """BASELINE.json configs[4] sweeps 2^16..2^23: above 2^20 complex points the engine runs a radix-2/4/8 split in
    front of 2^20-point sub-transforms (deep pyramids: up to 14 levels, pyramid_tail_kernel). Light input (noise + a few
    tones below the int8 wrap level) so the CPU oracle stays within seconds."""
    rng = np.random.default_rng(0xB200 + cfg.fft_size % 1013 + int(cfg.is_real))
    n = cfg.hop_samples
    scale = float(np.sqrt(2.0 ** 20 / cfg.fft_size))
    tones = [(rng.uniform(0.02, 0.48) if cfg.is_real else rng.uniform(-0.48, 0.48), rng.uniform(2e-4, 1.2e-3) * scale,
              rng.uniform(0, 2 * np.pi)) for _ in range(3)]
    pos = [0]

    def next_hop():
        t = np.arange(pos[0], pos[0] + n, dtype=np.float64)
        pos[0] += n
        if cfg.is_real:
            x = rng.standard_normal(n) * (1e-3 * scale)
            for f, a, ph in tones:
                x += a * np.cos(2 * np.pi * f * t + ph)
            return x.astype(np.float32)
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * (1e-3 * scale)
        for f, a, ph in tones:
            x += a * np.exp(1j * (2 * np.pi * f * t + ph))
        return x.astype(np.complex64)

    _run_frames_against_oracle(cfg, next_hop, 2)


def test_known_answer_tone_bin_and_level(gpu_required):
    """a complex tone at FFT bin k with amplitude A gives |X[k]|/N = A/2 (periodic Hann
    coherent gain 0.5), neighbours -A/4, at display index (k - N/2 - 1) mod N, with
    q = trunc(127 + 20 log10((A/2)^2) + 6.0206 log2 N) up to the polynomial's 0.05 dB error."""
    cfg = SpectrumConfig(sps=2_000_000, fft_size=1 << 16, is_real=False)
    N = cfg.fft_size
    k, A = 12345, 1e-3
    t = np.arange(3 * N // 2)
    x = (A * np.exp(2j * np.pi * k * t / N)).astype(np.complex64)
    eng = make_engine(cfg)
    bufs = [eng.malloc(cfg.hop_floats) for _ in range(2)]
    bufs[0][:] = hop_as_floats(x[: N // 2])
    bufs[1][:] = hop_as_floats(x[N // 2: N])
    eng.load_complex_input(bufs[0], bufs[1])
    eng.execute()
    X = eng.get_output_buffer().view(np.complex64)[:N]
    assert int(np.argmax(np.abs(X))) == k
    assert abs(abs(X[k]) - A / 2) < 1e-5 * A
    assert abs(X[k - 1] + A / 4 * np.exp(0j)) < 2e-5 * A and abs(X[k + 1] + A / 4) < 2e-5 * A
    q = eng.get_quantized_buffer()
    d = (k - N // 2 - 1) % N
    assert int(np.argmax(q[:N])) == d
    expect = 127 + 20 * np.log10((A / 2) ** 2) + 6.020599913 * np.log2(N)
    assert abs(int(q[d]) - expect) <= 1.0
    eng.close()
