"""A few forward launch groups of the bench configuration (2^20 c2c, 64 frames per launch), nothing else: the target of
`ncu --set full -k regex:... -s 6 -c 3` (the third group's pass 1, pass 2 and pyramid). Usage: fwd_once.py [groups] [real]"""
import sys
sys.path.insert(0, '.')
import torch
from phantomsdr_b200 import SpectrumConfig
from phantomsdr_b200.backend import B200FFT

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4
real = len(sys.argv) > 2 and sys.argv[2] == "real"
cfg = SpectrumConfig(sps=70_000_000, fft_size=1 << 21, is_real=True) if real else SpectrumConfig(sps=35_000_000, fft_size=1 << 20)
F = H = 64
eng = B200FFT(cfg.fft_size, 1, cfg.downsample_levels, 0, 0)
eng.set_output_additional_size(cfg.audio_fft_size)
eng.plan_r2c() if real else eng.plan_c2c()
eng.set_hop_ring(H)
eng.set_batch_frames(F)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
eng.set_stream(s.cuda_stream)
ring = torch.as_tensor(eng.device_hop_ring(H), device='cuda')
ring.normal_(0, 1e-3)
torch.cuda.synchronize()
for g in range(G):
    eng.execute_device(0, F)
torch.cuda.synchronize()
print("ok")
