"""Shared helpers for the parity tests: build matching oracle / engine pairs and drive them with
the reference's call order (src/fft.cpp:47-105)."""
from __future__ import annotations

import numpy as np

import oracle
from phantomsdr_b200 import SpectrumConfig
from phantomsdr_b200.synth import SignalSource, make_clients


def make_oracle_fft(cfg: SpectrumConfig) -> oracle.OracleFFT:
    f = oracle.OracleFFT(cfg.fft_size, cfg.downsample_levels, cfg.brightness_offset)
    f.set_output_additional_size(cfg.audio_fft_size)
    if cfg.is_real:
        f.plan_r2c()
    else:
        f.plan_c2c()
    return f


def make_engine(cfg: SpectrumConfig, device: int = 0):
    from phantomsdr_b200.backend import B200FFT

    e = B200FFT(cfg.fft_size, 1, cfg.downsample_levels, cfg.brightness_offset, device)
    e.set_output_additional_size(cfg.audio_fft_size)
    if cfg.is_real:
        e.plan_r2c()
    else:
        e.plan_c2c()
    return e


def hop_as_floats(hop: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(hop).view(np.float32)


def spectrum_tolerance_check(got: np.ndarray, ref: np.ndarray, tol: float = 1e-5):
    """max|delta| <= 1e-5 * max|X| per frame."""
    peak = float(np.abs(ref).max())
    err = float(np.abs(got.astype(np.complex128) - ref.astype(np.complex128)).max())
    assert err <= tol * peak, f"spectrum error {err:.3e} > {tol:g} * peak {peak:.3e}"
    return err / peak
