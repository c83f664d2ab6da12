"""Pins the oracle against the REFERENCE's own src/fft_impl.cpp (class FFTW: window, load_*_input, 1/N, vec_log2,
power_and_quantize, half_and_quantize pyramid) and src/signal.cpp (AudioClient::send_audio) compiled unmodified from
/root/reference into oracle/_ref/libphantom_ref_fft.so (oracle/ref_shim_fft.cpp + the stand-in headers of
oracle/ref_stub). FFTW3f itself is absent, so fftwf_execute runs the oracle's DFT on both sides: everything AROUND the
transforms is compared bit for bit - that is rows a3, a5-a7, a10, a12-a17 of SURVEY.md section 8.
Skipped only if the library was never built (needs /root/reference at build time)."""
import ctypes as C

import numpy as np
import pytest

import oracle
from oracle import AM, FM, LSB, USB

ref = oracle.ref_fft()
pytestmark = pytest.mark.skipif(ref is None, reason="oracle/_ref/libphantom_ref_fft.so not built (needs /root/reference)")
f32 = np.float32


def _view(ptr, n, dtype):
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n)


def _signal(rng, n, scale):
    """noise + a few strong/weak tones, in float32 (n scalar samples)."""
    t = np.arange(n)
    x = rng.standard_normal(n) * scale
    for k in range(4):
        x += scale * 10 ** rng.uniform(-1, 1.5) * np.cos(2 * np.pi * rng.uniform(0, 0.5) * t + k)
    return x.astype(f32)


# (fft_size, is_real, levels, brightness, additional): BASELINE configs 1, 2, 3 plus small and odd cases
FFT_CASES = [
    (1 << 10, False, 1, 0, 0),
    (1 << 12, True, 3, 2, 0),
    (1 << 17, False, 8, 0, 548),     # cfg 1: rtlsdr
    (1 << 20, False, 11, 0, 360),    # cfg 2: 35 MSPS IQ
    (1 << 21, True, 11, -1, 360),    # cfg 3: 70 MSPS real
]


@pytest.mark.parametrize("size,is_real,levels,bright,additional", FFT_CASES)
def test_fftw_class_bit_exact(size, is_real, levels, bright, additional):
    rng = np.random.default_rng(size + levels)
    h = ref.ref_fftw_create(size, levels, bright, additional, int(is_real))
    orc = oracle.OracleFFT(size, levels, bright)
    orc.set_output_additional_size(additional)
    (orc.plan_r2c if is_real else orc.plan_c2c)()
    hop = size // 2 * (1 if is_real else 2)
    R = size // 2 if is_real else size
    nin = size if is_real else 2 * size
    nout = size + 2 if is_real else 2 * size
    for frame in range(2):
        scale = [1e-3, 30.0][frame]  # second frame drives int8 values past 127 (the wrap region)
        a1, a2 = _signal(rng, hop, scale), _signal(rng, hop, scale)
        if is_real:
            ref.ref_fftw_load_real(h, a1.copy(), a2.copy())
            orc.load_real_input(a1, a2)
        else:
            ref.ref_fftw_load_complex(h, a1.copy(), a2.copy())
            orc.load_complex_input(a1.view(np.complex64), a2.view(np.complex64))
        assert np.array_equal(_view(ref.ref_fftw_input(h), nin, f32), orc.inbuf), "load_*_input differs"
        ref.ref_fftw_execute(h)
        orc.execute()
        got = _view(ref.ref_fftw_output(h), nout, f32)
        assert np.array_equal(got.view(np.uint32), orc.outbuf[:nout].view(np.uint32)), "normalised spectrum differs"
        nq = oracle.pyramid_size(R, levels)
        q_ref = _view(ref.ref_fftw_quantized(h), nq, np.int8)
        assert np.array_equal(q_ref, orc.quantized), "int8 pyramid differs"
    ref.ref_fftw_destroy(h)
    orc.close()


def test_quantiser_special_values_bit_exact():
    """All-zero, denormal-power and overflowing-power frames through the reference's own quantiser (finite inputs only:
    with NaN samples the reference's result depends on the NaN payload its vectorised loop happens to propagate)."""
    size, levels = 1 << 10, 4
    h = ref.ref_fftw_create(size, levels, 0, 0, 0)
    orc = oracle.OracleFFT(size, levels, 0)
    orc.plan_c2c()
    nq = oracle.pyramid_size(size, levels)
    seen = set()
    for value in (0.0, 1e-24, 3e-20, 1.0, 2e21, 3e38):
        a = np.zeros(size, f32)   # one hop = size/2 complex samples
        a[size // 2] = value      # a single real sample in the middle of the older half: the same |X| in every bin
        a[size // 2 + 3] = -value
        ref.ref_fftw_load_complex(h, a.copy(), a.copy())
        orc.load_complex_input(a.view(np.complex64), a.view(np.complex64))
        ref.ref_fftw_execute(h)
        orc.execute()
        q = _view(ref.ref_fftw_quantized(h), nq, np.int8)
        assert np.array_equal(q, orc.quantized), f"sample value {value}"
        seen.update(np.unique(q).tolist())
    assert -128 in seen and len(seen) > 20  # the floor, and the wrap region above 127, were both exercised
    ref.ref_fftw_destroy(h)
    orc.close()


def _spectrum(rng, R, extra, scale=1e-4):
    s = (rng.standard_normal(2 * (R + extra)) * scale).astype(f32)
    for k in rng.integers(0, R, 40):
        s[2 * k:2 * k + 2] += (rng.standard_normal(2) * scale * 300).astype(f32)
    s[2 * R:2 * (R + extra)] = s[:2 * extra]  # wrap tail, src/fft.cpp:96-97
    return s


@pytest.mark.parametrize("is_real", [False, True])
@pytest.mark.parametrize("n,sps", [(360, 12000), (548, 12000), (492, 12000)])
def test_send_audio_bit_exact(is_real, n, sps):
    """AudioClient::send_audio of the reference vs the oracle: PCM, power and the NaN drop, every mode, both parities
    of floor(mid), slices that hang over either edge of the placement window, mode and window changes mid-stream."""
    rng = np.random.default_rng(n + int(is_real))
    fft_size = 1 << 14
    R = fft_size // 2 if is_real else fft_size
    base = 0 if is_real else fft_size // 2 + 1
    cases = []
    for mode in (USB, LSB, AM, FM):
        for k in range(4):
            mid = float(rng.integers(n, R - n)) + [0.0, 0.25, 0.5, 0.99][k]
            m = int(np.floor(mid))
            if k == 0:
                l, r = (m, m + n // 3) if mode == USB else (m - n // 3, m) if mode == LSB else (m - n // 4, m + n // 4)
            elif k == 1:
                l, r = m - n // 2, m + n // 2          # full width, both sides
            elif k == 2:
                l, r = m + 5, m + 5 + n // 2            # window entirely above mid
            else:
                l, r = m - n + 10, m - 10               # window entirely below mid
            cases.append((mode, l, mid, r))
    clients = []
    for mode, l, mid, r in cases:
        a = ref.ref_audio_create(int(is_real), n, sps, R)
        ref.ref_audio_set_demodulation(a, mode)
        ref.ref_audio_set_range(a, l, mid, r)
        o = oracle.OracleClient(is_real, n, sps, R)
        o.set_audio_demodulation(mode)
        o.set_audio_range(l, mid, r)
        clients.append((a, o))
    names = {USB: b"USB", LSB: b"LSB", AM: b"AM", FM: b"FM"}
    nsent = 0
    for frame in range(34):
        spec = _spectrum(rng, R, n, scale=1e-4 * (1 + 20 * (frame % 7 == 3)))
        if frame == 20:
            spec[::97] = np.nan  # NaN guard, src/signal.cpp:266-271
        for ci, (a, o) in enumerate(clients):
            if frame == 24 and ci % 3 == 0:  # demodulation message: mode change + AGC reset, src/signal.cpp:316-328
                new_mode = (o.mode + 1) % 4
                ref.ref_audio_on_demodulation_message(a, names[new_mode])
                o.on_demodulation_message(new_mode)
            if frame == 27 and ci % 4 == 1:  # window message, src/signal.cpp:300-314
                nl, nm, nr = o.l + 3, o.mid + 3.5, o.r + 3
                ref.ref_audio_on_window_message(a, nl, nm, nr)
                assert o.on_window_message(nl, nm, nr)
            off = (o.l + base) % R
            pcm = np.zeros(n // 2, np.int32)
            pwr = np.zeros(1, f32)
            sent = ref.ref_audio_send(a, spec.ctypes.data + 8 * off, frame, pcm, pwr)
            ok, pcm_o, pwr_o, _ = o.send_audio(spec.view(np.complex64), fft_size, frame)
            assert bool(sent) == ok, f"frame {frame} client {ci}: sent {sent} vs oracle {ok}"
            if sent:
                nsent += 1
                assert np.array_equal(pcm, pcm_o), f"frame {frame} client {ci} mode {o.mode}: PCM differs"
                assert pwr[0] == f32(pwr_o) or (np.isnan(pwr[0]) and np.isnan(pwr_o)), \
                    f"frame {frame} client {ci}: pwr {pwr[0]} vs {pwr_o}"
    assert nsent > 30 * len(clients)
    for a, _ in clients:
        ref.ref_audio_destroy(a)
