"""Derived sizes of the spectrum server (host logic; reference src/spectrumserver.cpp:96-151,185-190
and src/fft.cpp:18,33). Pure Python integers/floats - no GPU needed."""
from __future__ import annotations

import math
from dataclasses import dataclass


def audio_fft_size(audio_max_sps: int, fft_size: int, sps: int) -> int:
    """src/spectrumserver.cpp:151: ceil((double)audio_max_sps * fft_size / sps / 4.) * 4"""
    return int(math.ceil(float(audio_max_sps) * fft_size / sps / 4.0) * 4)


def downsample_levels(fft_result_size: int, min_waterfall_fft: int = 1024) -> int:
    """src/spectrumserver.cpp:185-190"""
    levels, cur = 0, fft_result_size
    while cur >= min_waterfall_fft:
        levels += 1
        cur //= 2
    return levels


def skip_num(sps: int, fft_size: int) -> int:
    """src/fft.cpp:33: max(1, (int)floor(((float)sps / fft_size) / 10.) * 2); the quotient is a float32"""
    import numpy as np

    q = float(np.float32(sps) / np.float32(fft_size))
    return max(1, int(math.floor(q / 10.0)) * 2)


def level_offset(level: int, fft_result_size: int) -> int:
    """src/websocket.cpp:233: level i of the pyramid starts at sum_{j<i} (R >> j)"""
    return sum(fft_result_size >> j for j in range(level))


def pyramid_size(fft_result_size: int, levels: int) -> int:
    return level_offset(levels, fft_result_size)


@dataclass
class SpectrumConfig:
    """The [input] keys that size the hot path (src/spectrumserver.cpp:21-94)."""

    sps: int
    fft_size: int
    is_real: bool = False
    audio_sps: int = 12000
    waterfall_size: int = 1024
    brightness_offset: int = 0

    @property
    def fft_result_size(self) -> int:  # src/spectrumserver.cpp:99-105
        return self.fft_size // 2 if self.is_real else self.fft_size

    @property
    def base_idx(self) -> int:  # src/websocket.cpp:157-160
        return 0 if self.is_real else self.fft_size // 2 + 1

    @property
    def audio_fft_size(self) -> int:
        return audio_fft_size(self.audio_sps, self.fft_size, self.sps)

    @property
    def downsample_levels(self) -> int:
        return downsample_levels(self.fft_result_size, self.waterfall_size)

    @property
    def skip_num(self) -> int:
        return skip_num(self.sps, self.fft_size)

    @property
    def hop_samples(self) -> int:
        """new samples per frame (IQ pairs for c2c, reals for r2c), src/fft.cpp:49-67"""
        return self.fft_size // 2

    @property
    def hop_floats(self) -> int:
        """floats per input half-buffer, src/fft.cpp:18: fft_size / 2 * (2 - is_real)"""
        return self.fft_size // 2 * (2 - int(self.is_real))

    def slice_offset(self, l: int) -> int:  # src/websocket.cpp:182
        return (l + self.base_idx) % self.fft_result_size

    def passband_bins(self, hz: float) -> int:  # src/spectrumserver.cpp:127-129
        bw = self.sps / 2 if self.is_real else self.sps
        return int(hz * self.fft_result_size // bw)
