"""GPU: every kernel variant / scheduling option of the forward group must give bit-identical spectra and
pyramids on the same device-resident input (the variants only differ in where data is staged and which
kernel quantises), and the default path must survive a long saturated run (64-frame launch groups back to
back: the regime in which a barrier-protocol bug of the three-stage pass 2 once showed up)."""
import numpy as np
import pytest

from phantomsdr_b200 import SpectrumConfig
from helpers import make_engine

pytestmark = pytest.mark.gpu


def _device_engine(cfg, frames, ring):
    import torch

    from phantomsdr_b200 import backend as B

    eng = make_engine(cfg)
    eng.set_hop_ring(ring)
    eng.set_batch_frames(frames)
    ring_t = torch.as_tensor(eng.device_hop_ring(ring), device="cuda")
    g = torch.Generator(device="cuda")
    g.manual_seed(1234)
    ring_t.normal_(0, 1e-3, generator=g)
    torch.cuda.synchronize()
    return eng, torch, B


def _snapshot(eng, torch, frames, hop0=0):
    eng.execute_device(hop0, frames)
    eng.sync()
    spec = torch.as_tensor(eng.device_spectrum(frames), device="cuda").clone()
    quant = torch.as_tensor(eng.device_quantized(frames), device="cuda").clone()
    return spec, quant


def test_forward_variants_bit_identical(gpu_required):
    cfg = SpectrumConfig(sps=35_000_000, fft_size=1 << 20, is_real=False)
    F = 8
    eng, torch, B = _device_engine(cfg, F, 12)
    defaults = {B.OPT_FUSED_PYRAMID: -1, B.OPT_TMA: 2, B.OPT_PACKED_MATH: 1, B.OPT_FWD_LANES: 1, B.OPT_FWD_SUB_FRAMES: 64,
                B.OPT_PASS1_ORDER: 0, B.OPT_PYRAMID_LAG: 2}
    variants = [
        {},
        {B.OPT_PACKED_MATH: 0},
        {B.OPT_FUSED_PYRAMID: 2},
        {B.OPT_FUSED_PYRAMID: 1, B.OPT_TMA: 1},
        {B.OPT_TMA: 1},
        {B.OPT_TMA: 0},
        {B.OPT_PASS1_ORDER: 1},
        {B.OPT_PASS1_ORDER: 2},
        {B.OPT_FWD_SUB_FRAMES: 2},
        {B.OPT_FWD_LANES: 2, B.OPT_FWD_SUB_FRAMES: 2},
        {B.OPT_FWD_LANES: 4, B.OPT_FWD_SUB_FRAMES: 1},
        {B.OPT_TMA: 3},                                # pyramid fused into pass 2 (per-frame completion counters)
        {B.OPT_TMA: 3, B.OPT_PYRAMID_LAG: 1},
        {B.OPT_TMA: 3, B.OPT_PYRAMID_LAG: 3},
        {B.OPT_TMA: 3, B.OPT_PYRAMID_LAG: 8},          # lag >= frames: every block is produced by the tail loop
        {B.OPT_TMA: 3, B.OPT_FWD_LANES: 2, B.OPT_FWD_SUB_FRAMES: 2},
    ]
    ref = None
    for opts in variants:
        for k, v in {**defaults, **opts}.items():
            eng.set_option(k, v)
        spec, quant = _snapshot(eng, torch, F, hop0=2)
        if ref is None:
            ref = (spec, quant)
            assert float(spec.abs().max()) > 0
            continue
        if opts.get(B.OPT_TMA, 2) == 0:
            # the generic kernels share the arithmetic but not the instruction order of the twiddle products
            err = float((spec - ref[0]).abs().max()) / float(ref[0].abs().max())
            assert err <= 1e-5, f"{opts}: spectrum differs by {err:.2e}"
            continue
        assert torch.equal(spec, ref[0]), f"{opts}: spectrum not bit-identical"
        assert torch.equal(quant, ref[1]), f"{opts}: pyramid not bit-identical"
    eng.close()


def test_forward_saturated_soak(gpu_required):
    cfg = SpectrumConfig(sps=35_000_000, fft_size=1 << 20, is_real=False)
    F = 64
    eng, torch, B = _device_engine(cfg, F, F)
    first = _snapshot(eng, torch, F)
    for _ in range(150):  # ~0.1 s of back-to-back 64-frame launch groups, Y streamed through DRAM
        eng.execute_device(0, F)
    last = _snapshot(eng, torch, F)
    assert torch.equal(first[0], last[0]) and torch.equal(first[1], last[1])
    eng.close()
