"""Deterministic synthetic IQ / real input and client tables for tests and bench.py.

Generator: complex white Gaussian noise sigma = 1e-3 per component plus K tones at random display
positions (fractional bins allowed), amplitudes log-uniform in [1e-4, 1.5e-3] scaled by
sqrt(2^20 / N) so every level stays below the int8 wrap threshold (A < 2/sqrt(N)); 25 % of the
tones AM (1 kHz, m = 0.5), 25 % FM (+-2.5 kHz deviation, 400 Hz), the rest CW.
Seeds: numpy PCG64(0x5EED + config index).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np

from .sizes import SpectrumConfig

USB, LSB, AM, FM = 0, 1, 2, 3


@dataclass
class Tone:
    freq: float  # cycles per sample (IQ: in [-0.5, 0.5); real: in [0, 0.5))
    amp: float
    kind: int  # 0 CW, 1 AM, 2 FM
    phase: float


class SignalSource:
    """Streams hops of `cfg.hop_samples` samples: complex64 (IQ) or float32 (real)."""

    def __init__(self, cfg: SpectrumConfig, seed: int = 0x5EED, ntones: int = 32, noise_sigma: float = 1e-3):
        self.cfg = cfg
        self.rng = np.random.Generator(np.random.PCG64(seed))
        self.t0 = 0
        scale = float(np.sqrt(2.0 ** 20 / cfg.fft_size))
        self.noise_sigma = noise_sigma * min(1.0, scale)
        self.tones: List[Tone] = []
        for i in range(ntones):
            if cfg.is_real:
                f = self.rng.uniform(0.01, 0.49)
            else:
                f = self.rng.uniform(-0.49, 0.49)
            amp = float(np.exp(self.rng.uniform(np.log(1e-4), np.log(1.5e-3)))) * min(1.0, scale)
            kind = 1 if i % 4 == 1 else (2 if i % 4 == 3 else 0)
            self.tones.append(Tone(f, amp, kind, self.rng.uniform(0, 2 * np.pi)))

    def next_hop(self) -> np.ndarray:
        cfg = self.cfg
        n = cfg.hop_samples
        t = (self.t0 + np.arange(n)).astype(np.float64)
        self.t0 += n
        sps = float(cfg.sps)
        if cfg.is_real:
            x = self.rng.standard_normal(n) * self.noise_sigma
        else:
            x = (self.rng.standard_normal(n) + 1j * self.rng.standard_normal(n)) * self.noise_sigma
        for tn in self.tones:
            ph = 2 * np.pi * tn.freq * t + tn.phase
            a = tn.amp
            if tn.kind == 1:
                env = 1.0 + 0.5 * np.cos(2 * np.pi * 1000.0 / sps * t)
                s = a * env * (np.cos(ph) if cfg.is_real else np.exp(1j * ph))
            elif tn.kind == 2:
                beta = 2500.0 / 400.0
                ph = ph + beta * np.sin(2 * np.pi * 400.0 / sps * t)
                s = a * (np.cos(ph) if cfg.is_real else np.exp(1j * ph))
            else:
                s = a * (np.cos(ph) if cfg.is_real else np.exp(1j * ph))
            x = x + s
        return x.astype(np.float32) if cfg.is_real else x.astype(np.complex64)

    def display_bin(self, tone: Tone) -> float:
        """Display-axis position of a tone. IQ: FFT bin k <-> display (k - N/2 - 1) mod N."""
        cfg = self.cfg
        if cfg.is_real:
            return tone.freq * cfg.fft_size
        k = (tone.freq * cfg.fft_size) % cfg.fft_size
        return (k - cfg.fft_size // 2 - 1) % cfg.fft_size


@dataclass
class ClientSpec:
    l: int
    mid: float
    r: int
    mode: int


def make_clients(cfg: SpectrumConfig, count: int, seed: int = 0x5EED + 1, modes=(AM, USB, LSB), tones=None,
                 on_tone_fraction: float = 0.5) -> List[ClientSpec]:
    """Client table as the frontend would request it (html-svelte App.svelte:108-116, audio.js:276-283):
    USB [mid, mid+3 kHz], LSB [mid-3 kHz, mid], AM/FM mid +- 5 kHz; l = floor, r = ceil. Half of the
    clients sit on generated tones so the demodulators see signal, the rest see noise."""
    rng = np.random.Generator(np.random.PCG64(seed))
    R, n = cfg.fft_result_size, cfg.audio_fft_size
    o3, o5 = cfg.passband_bins(3000.0), cfg.passband_bins(5000.0)
    out: List[ClientSpec] = []
    for i in range(count):
        mode = modes[i % len(modes)]
        if tones and rng.uniform() < on_tone_fraction:
            mid = float(tones[int(rng.integers(len(tones)))]) + float(rng.uniform(-2.0, 2.0))
            mid = min(max(mid, n + 1.0), R - n - 2.0)
        else:
            mid = float(rng.uniform(n + 1, R - n - 2))
        if mode == USB:
            lo, hi = mid, mid + o3
        elif mode == LSB:
            lo, hi = mid - o3, mid
        else:
            lo, hi = mid - o5, mid + o5
        l, r = int(np.floor(lo)), int(np.ceil(hi))
        if r - l > n:
            r = l + n
        l = max(0, l)
        r = min(R - 1, r)
        out.append(ClientSpec(l, mid, r, mode))
    return out
