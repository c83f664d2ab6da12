"""Deterministic raw-sample inputs and client tables of the golden fixtures (shared by the generator
tests/golden/make_golden.py and by the tests). Pure integer arithmetic on the sample index: exact on
every platform."""
import numpy as np

from phantomsdr_b200 import SpectrumConfig, USB, LSB, AM, FM

CASES = {
    # BASELINE.json configs[0] shape: RTL-SDR u8 IQ, 2^17-point FFT at 2.88 MSPS (config.example.rtlsdr.toml:7) -> n = 548
    "rtlsdr_u8_iq_2p17": dict(cfg=SpectrumConfig(sps=2_880_000, fft_size=1 << 17, is_real=False), dtype="uint8", frames=5,
                              tones=[(12345, 30.0), (40000, 9.0), (-31000, 17.0), (70, 3.0)], seed=0x5EED),
    # real-input shape (RX888-like int16), 2^17-point r2c
    "real_s16_2p17": dict(cfg=SpectrumConfig(sps=4_370_000, fft_size=1 << 17, is_real=True), dtype="int16", frames=5,
                          tones=[(9000, 900.0), (22222, 300.0), (50001, 2500.0)], seed=0x5EED + 2),
}


def _hash32(k: np.ndarray, seed: int) -> np.ndarray:
    x = (k.astype(np.uint64) + np.uint64(seed)) & np.uint64(0xFFFFFFFF)
    x = (x ^ (x >> np.uint64(16))) * np.uint64(0x45D9F3B) & np.uint64(0xFFFFFFFF)
    x = (x ^ (x >> np.uint64(16))) * np.uint64(0x45D9F3B) & np.uint64(0xFFFFFFFF)
    x = x ^ (x >> np.uint64(16))
    return x


def raw_hop(case, hop_index: int) -> np.ndarray:
    """Hop `hop_index` as raw samples (interleaved I,Q for IQ). Noise = sum of two hashed uniforms; tones have
    integer bin numbers (cycles per fft_size samples) and amplitudes in LSB."""
    cfg = case["cfg"]
    ns = cfg.hop_samples
    t = np.arange(hop_index * ns, (hop_index + 1) * ns, dtype=np.int64)
    dt = np.dtype(case["dtype"])
    full = 256 if dt.itemsize == 1 else 65536
    spread = 16 if dt.itemsize == 1 else 1024

    def noise(stream):
        h = _hash32(t * 4 + stream, case["seed"])
        return ((h & np.uint64(0xFFFF)).astype(np.int64) % spread + ((h >> np.uint64(16)).astype(np.int64) % spread)
                - (spread - 1)).astype(np.float64)

    N = cfg.fft_size
    if cfg.is_real:
        x = noise(0)
        for k, a in case["tones"]:
            x += a * np.cos(2 * np.pi * ((k * t) % N) / N)
        v = np.rint(x).astype(np.int64)
    else:
        xi, xq = noise(0), noise(1)
        for k, a in case["tones"]:
            ph = 2 * np.pi * ((k * t) % N) / N
            xi += a * np.cos(ph)
            xq += a * np.sin(ph)
        v = np.empty(2 * ns, np.int64)
        v[0::2] = np.rint(xi)
        v[1::2] = np.rint(xq)
    if dt.kind == "u":
        v = v + full // 2
        return np.clip(v, 0, full - 1).astype(dt)
    return np.clip(v, -full // 2, full // 2 - 1).astype(dt)


def client_table(case):
    """(l, mid, r, mode) per client: one of each mode on every tone plus a few on noise."""
    cfg = case["cfg"]
    N, R, n = cfg.fft_size, cfg.fft_result_size, cfg.audio_fft_size
    o3, o5 = cfg.passband_bins(3000.0), cfg.passband_bins(5000.0)
    out = []
    for i, (k, _a) in enumerate(case["tones"]):
        d = float(k) if cfg.is_real else float((k % N - N // 2 - 1) % N)
        d = min(max(d, n + 2.0), R - n - 3.0)
        mode = (USB, LSB, AM, FM)[i % 4]
        frac = (0.0, 0.25, 0.5, 0.75)[i % 4]
        if mode == USB:
            mid = d - 10 + frac
            l, r = int(np.floor(mid)), int(np.ceil(mid + o3))
        elif mode == LSB:
            mid = d + 10 + frac
            l, r = int(np.floor(mid - o3)), int(np.ceil(mid))
        else:
            mid = d + 3 + frac
            l, r = int(np.floor(mid - o5)), int(np.ceil(mid + o5))
        r = min(r, l + n)
        out.append((l, mid, r, mode))
    out.append((R // 3, R // 3 + 0.5, R // 3 + min(o3, n), USB))
    out.append((R // 5, R // 5 + o5 + 0.125, min(R // 5 + 2 * o5, R // 5 + n), AM))
    return out
