#!/usr/bin/env python
"""Generates the committed golden fixtures tests/golden/*.npz from the CPU oracle (oracle/).

The reference ships no golden vectors for this path, and its FFTW build cannot run
here, so these fixtures are ORACLE-generated pins: they freeze the oracle's results (whose
helper stages are themselves pinned bit-for-bit against the reference's compiled sources by
tests/test_oracle_vs_ref.py) so that neither the oracle nor the CUDA path can drift unnoticed.

Inputs are raw u8 / s16 samples produced by an integer hash of the sample index (exact on every
platform, no RNG state), plus integer-parameter tones - the shape of BASELINE.json configs[0]
(u8 IQ from an RTL-SDR on stdin).

Run from the repo root:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import oracle  # noqa: E402
from golden_input import CASES, raw_hop, client_table  # noqa: E402


def run_case(name, case):
    cfg = case["cfg"]
    n = cfg.audio_fft_size
    R = cfg.fft_result_size
    orc = oracle.OracleFFT(cfg.fft_size, cfg.downsample_levels, cfg.brightness_offset)
    orc.set_output_additional_size(n)
    orc.plan_r2c() if cfg.is_real else orc.plan_c2c()
    clients = []
    specs = client_table(case)
    for (l, mid, r, mode) in specs:
        c = oracle.OracleClient(cfg.is_real, n, cfg.audio_sps, R)
        assert c.on_window_message(l, mid, r), (l, mid, r)
        c.set_audio_demodulation(mode)
        clients.append(c)
    nframes = case["frames"]
    rng = np.random.default_rng(1)  # only chooses WHICH bins are sampled into the fixture
    spec_idx = np.sort(rng.choice(R, size=384, replace=False)).astype(np.int64)
    spec_val = np.zeros((nframes, spec_idx.size), np.complex64)
    peak = np.zeros(nframes, np.float32)
    lvl3 = oracle.level_offset(3, R)
    q_hi = np.zeros((nframes, oracle.pyramid_size(R, cfg.downsample_levels) - lvl3), np.int8)
    q0_idx = np.sort(rng.choice(R, size=2048, replace=False)).astype(np.int64)
    q0_val = np.zeros((nframes, q0_idx.size), np.int8)
    pcm = np.zeros((nframes, len(specs), n // 2), np.int32)
    pwr = np.zeros((nframes, len(specs)), np.float32)
    for f in range(nframes):
        a1 = oracle.convert(raw_hop(case, f))
        a2 = oracle.convert(raw_hop(case, f + 1))
        (orc.load_real_input if cfg.is_real else orc.load_complex_input)(a1, a2)
        orc.execute()
        orc.wrap_copy(n)
        spec = orc.spectrum
        spec_val[f] = spec[spec_idx]
        peak[f] = np.abs(spec[:R]).max()
        q = orc.quantized
        q_hi[f] = q[lvl3:]
        q0_val[f] = q[q0_idx]
        for i, c in enumerate(clients):
            ok, p, pw, _ = c.send_audio(spec, cfg.fft_size, f)
            assert ok
            pcm[f, i] = p
            pwr[f, i] = pw
    out = Path(__file__).resolve().parent / f"{name}.npz"
    np.savez_compressed(out, spec_idx=spec_idx, spec_val=spec_val, peak=peak, q_hi=q_hi, q0_idx=q0_idx, q0_val=q0_val,
                        pcm=pcm, pwr=pwr)
    print(out, out.stat().st_size, "bytes")


if __name__ == "__main__":
    oracle.build()
    for name, case in CASES.items():
        run_case(name, case)
