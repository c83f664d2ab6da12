"""phantomsdr_b200 - B200-native spectrum / channeliser engine behind PhantomSDR's FFT-backend
interface (reference src/fft.h) and its send_audio / send_waterfall slots.

The compute path is hand-written sm_100a CUDA in ``csrc/`` behind the C ABI declared in
``include/phantomsdr_b200.h``; this package is the thin host-side mirror of the reference's
interface used by tests, bench.py and Python callers. There is no CPU fallback: importing
``phantomsdr_b200.backend`` without the built library, or creating an engine without a GPU, fails.
"""
from .sizes import SpectrumConfig, audio_fft_size, downsample_levels, skip_num, level_offset, pyramid_size  # noqa: F401

USB, LSB, AM, FM = 0, 1, 2, 3  # enum demodulation_mode, reference src/client.h:43
__all__ = ["SpectrumConfig", "audio_fft_size", "downsample_levels", "skip_num", "level_offset", "pyramid_size",
           "USB", "LSB", "AM", "FM"]
