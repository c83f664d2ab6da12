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
                B.OPT_PASS1_ORDER: 0, B.OPT_PYRAMID_LAG: 2, B.OPT_PASS1_SPLIT: 16, B.OPT_FWD_SMS: 0, B.OPT_STREAM_LAG1: 2,
                B.OPT_STREAM_LAG2: 4, B.OPT_STREAM_RING: 5}
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
        {B.OPT_PASS1_SPLIT: 2},                        # coarser pass-1 work units (two CTAs per column tile)
        {B.OPT_PASS1_SPLIT: 5},
        {B.OPT_FWD_SMS: 100},                          # pass-2 grid sized for fewer SMs
        {B.OPT_PACKED_MATH: 3},                        # table-driven quantiser for levels 0..2
        {B.OPT_TMA: 4},                                # all three stages in one persistent dataflow-scheduled launch
        {B.OPT_TMA: 4, B.OPT_STREAM_LAG1: 1, B.OPT_STREAM_LAG2: 2, B.OPT_STREAM_RING: 3},
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
        if opts.get(B.OPT_TMA, 2) in (0, 4):
            # the generic and the stream kernels share the arithmetic but not the instruction order / contraction of the
            # twiddle products: spectrum to rounding, int8 values off by one only where a power sits on a quantiser step
            err = float((spec - ref[0]).abs().max()) / float(ref[0].abs().max())
            assert err <= 1e-5, f"{opts}: spectrum differs by {err:.2e}"
            dq = (quant.to(torch.int16) - ref[1].to(torch.int16)).abs()
            dq = torch.minimum(dq, 256 - dq)  # (the reference's int8 wrap)
            assert int(dq.max()) <= 1 and float((dq != 0).float().mean()) <= 1e-4, f"{opts}: pyramid differs"
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


def test_chunked_demodulation_equals_sequential(gpu_required):
    """The frame-chunked warp-level demodulation (default) against the sequential one-CTA-per-client kernel: PCM, pwr and
    valid flags bit-identical for every mode, for chunk sizes that do and do not divide the batch, across batches, across a
    demodulation switch, and through frames dropped by the NaN guard (src/signal.cpp:266-275: the replay path)."""
    import torch

    from phantomsdr_b200 import USB, LSB, AM, FM
    from phantomsdr_b200 import backend as B
    from phantomsdr_b200.synth import SignalSource, make_clients
    from helpers import hop_as_floats

    cfg = SpectrumConfig(sps=4_370_000, fft_size=1 << 17)
    n, h = cfg.audio_fft_size, cfg.audio_fft_size // 2
    F, nblocks, nc = 16, 4, 41
    src = SignalSource(cfg, seed=44)
    specs = make_clients(cfg, nc, modes=(USB, LSB, AM, FM), tones=[src.display_bin(t) for t in src.tones])
    hops = np.stack([hop_as_floats(src.next_hop()) for _ in range(F * nblocks + 1)])
    hops[F + 5, 1000:1010] = np.nan   # frames F+4 and F+5 of the stream are dropped for every client
    results = []
    for chunk in (0, 8, 5, 1, 64):
        e = make_engine(cfg)
        e.set_hop_ring(F * nblocks + 1)
        e.set_batch_frames(F)
        e.set_option(B.OPT_DEMOD_CHUNK, chunk)
        e.clients_create(nc + 1, n, cfg.audio_sps)
        for i, c in enumerate(specs):
            e.client_open(i, c.l, c.mid, c.r, c.mode)
        ring = torch.as_tensor(e.device_hop_ring(F * nblocks + 1), device="cuda")
        ring.copy_(torch.from_numpy(hops))
        torch.cuda.synchronize()
        out = []
        for k in range(nblocks):
            if k == 2:
                e.client_set_demodulation(2, FM)
                e.client_set_demodulation(7, USB)
            e.execute_device(k * F, F)
            e.clients_execute_device(k * F, F)
            for f in range(F):
                pcm, pwr, valid = e.clients_fetch(f)
                out.append((pcm.copy(), pwr.copy(), valid.copy()))
        results.append(out)
        e.close()
    ref = results[0]
    # frames F+4 and F+5 contain the NaN hop; AM / FM also lose F+6 (its overlap half is the NaN frame's upper half) and FM
    # F+7 (the discriminator's first step starts from the last sample of F+6)
    assert not ref[F + 4][2][:nc].any() and not ref[F + 5][2][:nc].any() and ref[F + 8][2][:nc].all(), "NaN frames not dropped"
    for chunk, res in zip((8, 5, 1, 64), results[1:]):
        for f, ((pa, wa, va), (pb, wb, vb)) in enumerate(zip(ref, res)):
            assert np.array_equal(va, vb), f"chunk {chunk} frame {f}: valid flags differ"
            assert np.array_equal(pa, pb), f"chunk {chunk} frame {f}: PCM differs at clients {np.flatnonzero((pa != pb).any(axis=1))[:8]}"
            assert np.array_equal(wa.view(np.uint32), wb.view(np.uint32)), f"chunk {chunk} frame {f}: pwr differs"
    assert any(p.any() for p, _, _ in ref[3 * F:]), "AGC never opened: test is vacuous"


@pytest.mark.parametrize("is_real", [False, True])
def test_raw_sample_formats_on_the_tma_path(gpu_required, is_real):
    """2^20-point transforms read raw ADC samples through the TMA pass 1 (SampleConverter fused): the
    result must be bit-identical to the float path fed with the converted samples ((x ^ topbit) / 2^(bits-1),
    src/samplereader.cpp:29-40,59-66)."""
    import torch

    from phantomsdr_b200 import backend as B

    cfg = SpectrumConfig(sps=70_000_000, fft_size=1 << 21, is_real=True) if is_real else SpectrumConfig(sps=35_000_000, fft_size=1 << 20)
    F, H = 4, 6
    g = torch.Generator(device="cuda")
    g.manual_seed(77)
    per_hop = cfg.hop_floats  # scalar samples per hop
    cases = [(B.FMT_U8, torch.uint8, 8), (B.FMT_S8, torch.int8, 8), (B.FMT_U16, torch.int16, 16), (B.FMT_S16, torch.int16, 16)]
    for fmt, tdt, bits in cases:
        raw = torch.randint(0, 1 << bits, (H * per_hop,), generator=g, device="cuda", dtype=torch.int32)
        # the float the reference's converter makes of each raw value
        unsigned = fmt in (B.FMT_U8, B.FMT_U16)
        signed = (raw ^ (1 << (bits - 1))) if unsigned else raw
        signed = torch.where(signed >= (1 << (bits - 1)), signed - (1 << bits), signed)
        as_float = signed.to(torch.float32) / float(1 << (bits - 1))
        out = []
        for use_raw in (True, False):
            eng = make_engine(cfg)
            eng.set_option(B.OPT_INPUT_FORMAT, fmt if use_raw else B.FMT_F32)
            eng.set_hop_ring(H)
            eng.set_batch_frames(F)
            ring = torch.as_tensor(eng.device_hop_ring(H), device="cuda").reshape(-1)  # (allocated for float samples)
            if use_raw:  # raw hops are packed back to back from the start of the ring
                view = ring.view(torch.uint8) if bits == 8 else ring.view(torch.int16)
                view[: H * per_hop] = raw.to(torch.uint8) if bits == 8 else raw.to(torch.int16)
            else:
                ring[: H * per_hop] = as_float
            torch.cuda.synchronize()
            out.append(_snapshot(eng, torch, F, hop0=1))
            eng.close()
        assert float(out[0][0].abs().max()) > 0
        assert torch.equal(out[0][0], out[1][0]), f"format {fmt}: spectrum differs from the float path"
        assert torch.equal(out[0][1], out[1][1]), f"format {fmt}: pyramid differs from the float path"


@pytest.mark.parametrize("log2n", [17, 21])
def test_r2c_split_kernel_equals_fused_split(gpu_required, log2n):
    """r2c: the Hermitian split as a kernel of its own followed by the c2c pyramid kernel must give exactly the
    spectrum (Nyquist bin included) and the pyramid of the one-kernel split-and-quantise path, with and without a
    waterfall cadence."""
    cfg = SpectrumConfig(sps=70_000_000 >> (21 - log2n), fft_size=1 << log2n, is_real=True)
    F = 6
    eng, torch, B = _device_engine(cfg, F, F + 2)
    for skip in (1, 4):
        eng.set_waterfall_cadence(skip)
        res = []
        for split in (0, 1):
            eng.set_option(B.OPT_R2C_SPLIT_KERNEL, split)
            eng.set_frame_number(2)
            torch.as_tensor(eng.device_quantized(F), device="cuda").zero_()
            res.append(_snapshot(eng, torch, F, hop0=1))
        assert float(res[0][0].abs().max()) > 0
        assert torch.equal(res[0][0], res[1][0]), f"cadence {skip}: spectrum differs"
        assert torch.equal(res[0][1], res[1][1]), f"cadence {skip}: pyramid differs"
        assert int((res[1][1] != 0).sum()) > 0
    eng.close()
