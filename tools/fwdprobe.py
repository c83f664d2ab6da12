"""Times the forward stages alone (HBM-resident ring) for a few engine options. Usage: python tools/fwdprobe.py [batch]"""
import sys, json
sys.path.insert(0, '.')
import torch
from phantomsdr_b200 import SpectrumConfig
from phantomsdr_b200.backend import B200FFT, OPT_STAGE_MASK, OPT_FUSED_PYRAMID, OPT_TMA

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


for name, opts in (("tma power(2)", {}), ("tma fused(1)", {OPT_FUSED_PYRAMID: 1}), ("tma unfused(0)", {OPT_FUSED_PYRAMID: 0}),
                   ("no tma power", {OPT_TMA: 0})):
    eng.set_option(OPT_FUSED_PYRAMID, 2)
    eng.set_option(OPT_TMA, 1)
    for k, v in opts.items():
        eng.set_option(k, v)
    print(f"{name:14s} batch {F}: pass1 {t(1):.2f}  pass2 {t(2):.2f}  pyramid {t(4):.2f}  all {t(7):.2f} us/frame")
