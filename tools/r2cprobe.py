"""cfg 3 forward group (2^21 r2c): us per frame per stage. Usage: r2cprobe.py [batch]"""
import sys
sys.path.insert(0, '.')
import torch
from phantomsdr_b200 import SpectrumConfig
from phantomsdr_b200.backend import B200FFT, OPT_STAGE_MASK

F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = SpectrumConfig(sps=70_000_000, fft_size=1 << 21, is_real=True)
eng = B200FFT(cfg.fft_size, 1, cfg.downsample_levels, 0, 0)
eng.set_output_additional_size(cfg.audio_fft_size)
eng.plan_r2c()
eng.set_hop_ring(F)
eng.set_batch_frames(F)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
eng.set_stream(s.cuda_stream)
ring = torch.as_tensor(eng.device_hop_ring(F), device='cuda')
ring.normal_(0, 1e-3)
torch.cuda.synchronize()
out = []
for mask in (1, 2, 4, 7):
    eng.set_option(OPT_STAGE_MASK, mask)
    for _ in range(2):
        eng.execute_device(0, F)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(10):
        eng.execute_device(0, F)
    b.record(s)
    torch.cuda.synchronize()
    out.append(a.elapsed_time(b) * 1e3 / (10 * F))
print(f"r2c 2^21: pass1 {out[0]:.2f}  pass2 {out[1]:.2f}  split+pyramid {out[2]:.2f}  all {out[3]:.2f} us/frame")
