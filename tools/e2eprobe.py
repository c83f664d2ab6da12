"""Where a host block spends its time: b200_submit_block / b200_wait_block with parts of the work left out.
Usage: e2eprobe.py [u8|s16|f32]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from phantomsdr_b200 import SpectrumConfig, AM, USB, LSB
from phantomsdr_b200.backend import B200FFT, OPT_INPUT_FORMAT, OPT_PCM16, FMT_F32, FMT_U8, FMT_S16
from phantomsdr_b200.synth import make_clients

fmt_name = sys.argv[1] if len(sys.argv) > 1 else "u8"
fmt, dt = {"u8": (FMT_U8, np.uint8), "s16": (FMT_S16, np.int16), "f32": (FMT_F32, np.float32)}[fmt_name]
cfg = SpectrumConfig(sps=35_000_000, fft_size=1 << 20)
n, h, F, NC, DEPTH, NB = cfg.audio_fft_size, cfg.audio_fft_size // 2, 64, 1024, 4, 12


def run(label, clients=True, pyr=True, skip=6, pcm=True):
    eng = B200FFT(cfg.fft_size, 1, cfg.downsample_levels, 0, 0)
    eng.set_output_additional_size(n)
    eng.plan_c2c()
    eng.set_option(OPT_INPUT_FORMAT, fmt)
    eng.set_hop_ring(DEPTH * F + 2)
    eng.set_batch_frames(F)
    eng.set_pipeline(2)
    eng.set_option(OPT_PCM16, 1)
    eng.set_waterfall_cadence(skip)
    if clients:
        eng.clients_create(NC, n, 12000)
        for i, c in enumerate(make_clients(cfg, NC, modes=(AM, USB, LSB))):
            eng.client_open(i, c.l, c.mid, c.r, c.mode)
    contiguous = "--separate" not in sys.argv
    if contiguous:
        buf = eng.pinned(F * cfg.hop_floats * np.dtype(dt).itemsize, dt)
        halves = [buf[k * cfg.hop_floats:(k + 1) * cfg.hop_floats] for k in range(F)]
    else:
        halves = [eng.pinned(cfg.hop_floats * np.dtype(dt).itemsize, dt) for _ in range(F)]
    for hb in halves:
        hb[:] = 1
    sets = [dict(pcm=eng.pinned(2 * F * NC * h, np.uint8) if (clients and pcm) else None,
                 pwr=eng.pinned(4 * F * NC, np.float32) if clients else None,
                 valid=eng.pinned(F * NC, np.uint8) if clients else None,
                 pyr=eng.pinned(F * eng.pyramid_bytes, np.int8) if pyr else None) for _ in range(DEPTH)]

    def go(blocks, f0):
        eng.stream_prime(halves[0])
        for k in range(blocks):
            st = sets[k % DEPTH]
            if k >= DEPTH:
                eng.wait_block()
            eng.submit_block(halves, f0 + k * F, st["pcm"], st["pwr"], st["valid"], st["pyr"])
        for _ in range(min(DEPTH, blocks)):
            eng.wait_block()

    go(3, 0)
    t0 = time.perf_counter()
    go(NB, 3 * F)
    dt_s = time.perf_counter() - t0
    print(f"{fmt_name} {label:44s} {1e3 * dt_s / NB:6.2f} ms per block of {F} frames = {NB * F * cfg.hop_samples / dt_s / 1e6:8.0f} MS/s", flush=True)
    eng.close()


run("everything (pyramid every 6th frame, int16 PCM)")
run("no pyramid copy", pyr=False)
run("no clients", clients=False)
run("no clients, no pyramid copy", clients=False, pyr=False)
run("clients, PCM not copied", pcm=False)
run("pyramid of every frame", skip=1)
