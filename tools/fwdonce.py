"""One forward launch group after warm-up (for ncu captures). Usage: python tools/fwdonce.py [batch] [groups]"""
import sys
sys.path.insert(0, '.')
import torch
from phantomsdr_b200 import SpectrumConfig
from phantomsdr_b200.backend import B200FFT

F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
G = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H = 64
cfg = SpectrumConfig(sps=35_000_000, fft_size=1 << 20)
eng = B200FFT(cfg.fft_size, 1, cfg.downsample_levels, 0, 0)
eng.set_output_additional_size(cfg.audio_fft_size)
eng.plan_c2c()
eng.set_hop_ring(H)
eng.set_batch_frames(F)
ring = torch.as_tensor(eng.device_hop_ring(H), device='cuda')
ring.normal_(0, 1e-3)
for g in range(G):
    eng.execute_device((g * F) % H, F)
eng.sync()
print("done")
