"""GPU: the pipelined host-block API (b200_stream_prime / b200_submit_block / b200_wait_block) must give exactly
what the per-frame reference-shaped calls give (load_complex_input -> execute -> clients_execute) on the same input."""
import numpy as np
import pytest

from phantomsdr_b200 import SpectrumConfig, USB, LSB, AM, FM
from phantomsdr_b200.synth import SignalSource, make_clients
from helpers import make_engine, hop_as_floats

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("is_real,depth", [(False, 2), (True, 2), (False, 3), (False, 4)])
def test_block_streaming_equals_per_frame_calls(gpu_required, is_real, depth):
    cfg = SpectrumConfig(sps=4_370_000 * (2 if is_real else 1), fft_size=(1 << 18) if is_real else (1 << 17), is_real=is_real)
    n, h = cfg.audio_fft_size, cfg.audio_fft_size // 2
    F, nblocks, nc = 4, 3 if depth == 2 else 7, 12   # (depth = blocks in flight = what the hop ring holds)
    src = SignalSource(cfg, seed=21)
    hops = [hop_as_floats(src.next_hop()).copy() for _ in range(F * nblocks + 1)]
    specs = make_clients(cfg, nc, modes=(USB, LSB, AM, FM), tones=[src.display_bin(t) for t in src.tones])

    def setup(e):
        e.clients_create(nc, n, cfg.audio_sps)
        for i, c in enumerate(specs):
            e.client_open(i, c.l, c.mid, c.r, c.mode)

    # reference-shaped per-frame calls
    a = make_engine(cfg)
    setup(a)
    ring = [a.malloc(cfg.hop_floats) for _ in range(3)]
    want_pcm, want_pyr = [], []
    ring[0][:] = hops[0]
    for f in range(F * nblocks):
        ring[(f + 1) % 3][:] = hops[f + 1]
        (a.load_real_input if is_real else a.load_complex_input)(ring[f % 3], ring[(f + 1) % 3])
        a.execute()
        pcm, pwr, valid = a.clients_execute(f)
        assert valid[:nc].all()
        want_pcm.append(pcm.copy())
        want_pyr.append(a.get_quantized_buffer().copy())
    a.close()

    # pipelined blocks
    b = make_engine(cfg)
    b.set_hop_ring(depth * F + 2)
    b.set_batch_frames(F)
    b.set_pipeline(2)
    setup(b)
    sets = []
    for _ in range(depth):
        if depth == 4:  # a block's halves in one host buffer: the engine merges them into one copy per contiguous run
            blockbuf = b.malloc(F * cfg.hop_floats)
            halves = [blockbuf[f * cfg.hop_floats:(f + 1) * cfg.hop_floats] for f in range(F)]
        else:
            halves = [b.malloc(cfg.hop_floats) for _ in range(F)]
        sets.append(dict(halves=halves, pcm=b.pinned(4 * F * nc * h, np.int32),
                         pwr=b.pinned(4 * F * nc, np.float32), valid=b.pinned(F * nc, np.uint8),
                         pyr=b.pinned(F * b.pyramid_bytes, np.int8)))
    prime = b.malloc(cfg.hop_floats)
    prime[:] = hops[0]
    b.stream_prime(prime)
    got_pcm, got_pyr = [], []

    def collect(st):
        got_pcm.extend(st["pcm"].reshape(F, nc, h).copy())
        got_pyr.extend(st["pyr"].reshape(F, -1).copy())

    for k in range(nblocks):
        st = sets[k % depth]
        if k >= depth:
            b.wait_block()
            collect(sets[k % depth])
        for f in range(F):
            st["halves"][f][:] = hops[1 + k * F + f]
        b.submit_block(st["halves"], k * F, st["pcm"], st["pwr"], st["valid"], st["pyr"])
    # drain in submission order
    pending = list(range(max(0, nblocks - depth), nblocks))
    for k in pending:
        b.wait_block()
        collect(sets[k % depth])
    b.close()
    assert len(got_pcm) == F * nblocks
    for f in range(F * nblocks):
        assert np.array_equal(got_pyr[f], want_pyr[f]), f"frame {f}: pyramid differs"
        assert np.array_equal(got_pcm[f], want_pcm[f][:nc]), f"frame {f}: PCM differs"


def test_pipelined_tail_equals_phase_tail(gpu_required):
    """The frame-skewed tail pipeline (default for batches >= 4 frames) must be bit-identical to the per-phase tail
    kernel, including AGC warm-up (zeros until the look-ahead fills), a mode switch (AGC reset) and batches that
    straddle the moment the look-ahead fills."""
    from phantomsdr_b200.backend import OPT_TAIL_PIPELINE
    import torch

    cfg = SpectrumConfig(sps=4_370_000, fft_size=1 << 17)
    n, h = cfg.audio_fft_size, cfg.audio_fft_size // 2
    F, nblocks, nc = 8, 5, 19
    src = SignalSource(cfg, seed=33)
    specs = make_clients(cfg, nc, modes=(USB, LSB, AM, FM), tones=[src.display_bin(t) for t in src.tones])
    hops = np.stack([hop_as_floats(src.next_hop()) for _ in range(F * nblocks + 1)])
    results = []
    for pipe in (0, 1):
        e = make_engine(cfg)
        e.set_hop_ring(F * nblocks + 1)
        e.set_batch_frames(F)
        e.set_option(OPT_TAIL_PIPELINE, pipe)
        e.clients_create(nc + 2, n, cfg.audio_sps)
        for i, c in enumerate(specs):
            e.client_open(i, c.l, c.mid, c.r, c.mode)
        ring = torch.as_tensor(e.device_hop_ring(F * nblocks + 1), device="cuda")
        ring.copy_(torch.from_numpy(hops))
        torch.cuda.synchronize()
        out = []
        for k in range(nblocks):
            if k == 2:
                e.client_set_demodulation(3, AM)   # AGC reset in the middle of the stream
                e.client_set_demodulation(4, USB)
            e.execute_device(k * F, F)
            e.clients_execute_device(k * F, F)
            for f in range(F):
                pcm, pwr, valid = e.clients_fetch(f)
                out.append((pcm.copy(), valid.copy()))
        results.append(out)
        e.close()
    for f, ((pa, va), (pb, vb)) in enumerate(zip(*results)):
        assert np.array_equal(va, vb), f"frame {f}: valid flags differ"
        assert np.array_equal(pa, pb), f"frame {f}: PCM differs at clients {np.flatnonzero((pa != pb).any(axis=1))[:8]}"
    assert any(p.any() for p, _ in results[1][20:]), "AGC never opened: test is vacuous"
