"""Client kernels alone: demodulation (sequential / frame-chunked variants) and tails, device-timed. Usage: cliprobe.py [clients] [batch]"""
import sys
sys.path.insert(0, '.')
import torch
from phantomsdr_b200 import SpectrumConfig, AM, USB, LSB
from phantomsdr_b200.backend import B200FFT, OPT_DEMOD_CHUNK, OPT_CLIENT_STAGE_MASK, OPT_DEMOD_GENERIC
from phantomsdr_b200.synth import make_clients

NC = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
F = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg = SpectrumConfig(sps=35_000_000, fft_size=1 << 20)
n = cfg.audio_fft_size
eng = B200FFT(cfg.fft_size, 1, cfg.downsample_levels, 0, 0)
eng.set_output_additional_size(n)
eng.plan_c2c()
eng.set_hop_ring(F + 1)
eng.set_batch_frames(F)
eng.clients_create(NC, n, 12000)
for i, c in enumerate(make_clients(cfg, NC, modes=(AM, USB, LSB))):
    eng.client_open(i, c.l, c.mid, c.r, c.mode)
s = torch.cuda.Stream()
torch.cuda.set_stream(s)
eng.set_stream(s.cuda_stream)
ring = torch.as_tensor(eng.device_hop_ring(F + 1), device='cuda')
ring.normal_(0, 1e-3)
eng.execute_device(0, F)
fn = 0


def t(reps=10):
    global fn
    for _ in range(2):
        eng.clients_execute_device(fn, F); fn += F
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(reps):
        eng.clients_execute_device(fn, F); fn += F
    b.record(s)
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / (reps * F)


for name, mask, chunk, gen in (("demod sequential", 1, 0, 0), ("demod chunk 8 run-time plan", 1, 8, 1), ("demod chunk 4", 1, 4, 0),
                               ("demod chunk 7", 1, 7, 0), ("demod chunk 8", 1, 8, 0), ("demod chunk 10", 1, 10, 0), ("demod chunk 11", 1, 11, 0),
                               ("demod chunk 16", 1, 16, 0), ("demod chunk 32", 1, 32, 0),
                               ("tails", 2, 8, 0), ("demod chunk 8 + tails", 3, 8, 0)):
    eng.set_option(OPT_CLIENT_STAGE_MASK, mask)
    eng.set_option(OPT_DEMOD_CHUNK, chunk)
    eng.set_option(OPT_DEMOD_GENERIC, gen)
    print(f"{NC} clients, batch {F}: {name:24s} {t():7.2f} us/frame", flush=True)
