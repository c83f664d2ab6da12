"""Forward-group variants: bit-for-bit comparison against the first variant, then per-stage timings (edit VARIANTS).
Usage: python tools/fwdprobe.py [batch] [name substrings...]"""
import sys
sys.path.insert(0, '.')
import torch
from phantomsdr_b200 import SpectrumConfig
from phantomsdr_b200.backend import (B200FFT, OPT_STAGE_MASK, OPT_FUSED_PYRAMID, OPT_TMA, OPT_PACKED_MATH, OPT_FWD_LANES,
                                     OPT_FWD_SUB_FRAMES, OPT_PASS1_ORDER, OPT_PYRAMID_LAG, OPT_STREAM_GRID, OPT_STREAM_LAG1,
                                     OPT_STREAM_LAG2, OPT_STREAM_RING, OPT_PASS1_SPLIT)

F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H = 64
cfg = SpectrumConfig(sps=35_000_000, fft_size=1 << 20)
eng = B200FFT(cfg.fft_size, 1, cfg.downsample_levels, 0, 0)
eng.set_output_additional_size(cfg.audio_fft_size)
eng.plan_c2c()
eng.set_hop_ring(H)
eng.set_batch_frames(F)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
eng.set_stream(s.cuda_stream)
ring = torch.as_tensor(eng.device_hop_ring(H), device='cuda')
ring.normal_(0, 1e-3)
torch.cuda.synchronize()

DEFAULTS = {OPT_FUSED_PYRAMID: -1, OPT_PYRAMID_LAG: 2, OPT_TMA: 2, OPT_PACKED_MATH: 1, OPT_FWD_LANES: 1, OPT_FWD_SUB_FRAMES: 64, OPT_PASS1_ORDER: 0,
            OPT_STREAM_GRID: 0, OPT_STREAM_LAG1: 2, OPT_STREAM_LAG2: 4, OPT_STREAM_RING: 5, OPT_PASS1_SPLIT: 16}


def configure(opts):
    for k, v in {**DEFAULTS, **opts}.items():
        eng.set_option(k, v)


def snapshot():
    eng.execute_device(3, F)
    torch.cuda.synchronize()
    spec = torch.as_tensor(eng.device_spectrum(F), device='cuda').clone()
    quant = torch.as_tensor(eng.device_quantized(F), device='cuda').clone()
    return spec, quant


def t(mask, reps=10):
    eng.set_option(OPT_STAGE_MASK, mask)
    for g in range(H // F):
        eng.execute_device(g * F, F)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(reps):
        for g in range(H // F):
            eng.execute_device(g * F, F)
    b.record(s)
    torch.cuda.synchronize()
    eng.set_option(OPT_STAGE_MASK, 7)
    return a.elapsed_time(b) * 1e3 / (reps * H)


VARIANTS = [
    ("default (tma2 fuse0)", {}),
    ("scalar quantiser", {OPT_PACKED_MATH: 0}),
    ("pass-1 split 2", {OPT_PASS1_SPLIT: 2}),
    ("tma1 two-CTA pass 2", {OPT_TMA: 1}),
    ("power plane (fuse 2)", {OPT_FUSED_PYRAMID: 2}),
    ("pass-1 order 1", {OPT_PASS1_ORDER: 1}),
    ("pass-1 order 2", {OPT_PASS1_ORDER: 2}),
    ("2 lanes x 4 frames", {OPT_FWD_LANES: 2, OPT_FWD_SUB_FRAMES: 4}),
    ("2 lanes x 32 frames", {OPT_FWD_LANES: 2, OPT_FWD_SUB_FRAMES: 32}),
    ("2 lanes x 16 frames", {OPT_FWD_LANES: 2, OPT_FWD_SUB_FRAMES: 16}),
    ("4 lanes x 16 frames", {OPT_FWD_LANES: 4, OPT_FWD_SUB_FRAMES: 16}),
    ("3 lanes x 8 frames", {OPT_FWD_LANES: 3, OPT_FWD_SUB_FRAMES: 8}),
    ("sub-batches of 8", {OPT_FWD_SUB_FRAMES: 8}),
    ("tma3 fused lag1", {OPT_TMA: 3, OPT_PYRAMID_LAG: 1}),
    ("tma3 fused lag2", {OPT_TMA: 3, OPT_PYRAMID_LAG: 2}),
    ("table quantiser", {OPT_PACKED_MATH: 3}),
    ("generic (tma 0)", {OPT_TMA: 0}),
    ("generic fuse 0", {OPT_TMA: 0, OPT_FUSED_PYRAMID: 0}),
    ("stream (tma 4)", {OPT_TMA: 4}),
    ("stream lag 1/2 ring 2", {OPT_TMA: 4, OPT_STREAM_LAG1: 1, OPT_STREAM_LAG2: 2, OPT_STREAM_RING: 2}),
    ("stream lag 1/2 ring 3", {OPT_TMA: 4, OPT_STREAM_LAG1: 1, OPT_STREAM_LAG2: 2, OPT_STREAM_RING: 3}),
    ("stream lag 2/3 ring 3", {OPT_TMA: 4, OPT_STREAM_LAG1: 2, OPT_STREAM_LAG2: 3, OPT_STREAM_RING: 3}),
]
if len(sys.argv) > 2:  # keep the reference variant plus those whose name contains one of the given substrings
    VARIANTS = [VARIANTS[0]] + [v for v in VARIANTS[1:] if any(k in v[0] for k in sys.argv[2:])]
ref = None
for name, opts in VARIANTS:
    if F < 16 and opts.get(OPT_FWD_SUB_FRAMES, 64) * opts.get(OPT_FWD_LANES, 1) > F:
        continue
    configure(opts)
    spec, quant = snapshot()
    if ref is None:
        ref = (spec, quant)
        chk = "reference"
    else:
        ds = (spec - ref[0]).abs().max().item()
        nq = (quant != ref[1]).sum().item()
        chk = f"spec maxdiff {ds:.3e} (max {ref[0].abs().max().item():.3e})  pyramid bytes differing {nq}"
        if opts.get(OPT_TMA, 2) == 4:
            # the stream kernel builds its window on the fly: its spectrum differs in the last bits, so its pyramid is checked
            # against the stand-alone quantiser run on the stream kernel's own spectrum (still in the bank)
            eng.set_option(OPT_TMA, 2)
            eng.set_option(OPT_STAGE_MASK, 4)
            eng.execute_device(3, F)
            torch.cuda.synchronize()
            q2 = torch.as_tensor(eng.device_quantized(F), device='cuda')
            chk += f" | vs quantiser on own spectrum: {(quant != q2).sum().item()} bytes differ"
            eng.set_option(OPT_STAGE_MASK, 7)
            eng.set_option(OPT_TMA, 4)
    full = opts.get(OPT_TMA, 2) != 3 and OPT_FWD_LANES not in opts and OPT_FWD_SUB_FRAMES not in opts
    if full:
        print(f"{name:20s} batch {F}: pass1 {t(1):.2f}  pass2 {t(2):.2f}  pyramid {t(4):.2f}  all {t(7):.2f} us/frame | {chk}", flush=True)
    else:
        print(f"{name:20s} batch {F}: all {t(7):.2f} us/frame | {chk}", flush=True)
