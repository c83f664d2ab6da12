"""GPU: the waterfall slot (b200_waterfall_gather = N x WaterfallClient::send_waterfall, src/waterfall.cpp:44-51
with the level pointer of src/websocket.cpp:227-233), the waterfall cadence (8f N3: pyramids only on the frames the
reference sends, src/fft.cpp:33,102-104) and the int16 PCM hand-off (8f N2)."""
import numpy as np
import pytest

import oracle
from phantomsdr_b200 import SpectrumConfig, USB, LSB, AM, FM
from phantomsdr_b200.synth import SignalSource, make_clients
from helpers import make_engine, make_oracle_fft, hop_as_floats

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("is_real", [False, True])
def test_waterfall_gather_matches_reference_slices(gpu_required, is_real):
    cfg = SpectrumConfig(sps=4_370_000 * (2 if is_real else 1), fft_size=(1 << 18) if is_real else (1 << 17), is_real=is_real)
    R, L = cfg.fft_result_size, cfg.downsample_levels
    src = SignalSource(cfg, seed=5)
    eng, orc = make_engine(cfg), make_oracle_fft(cfg)
    a1, a2 = eng.malloc(cfg.hop_floats), eng.malloc(cfg.hop_floats)
    a1[:] = hop_as_floats(src.next_hop())
    a2[:] = hop_as_floats(src.next_hop())
    (eng.load_real_input if is_real else eng.load_complex_input)(a1, a2)
    eng.execute()
    want = orc.requantize_from(eng.get_output_buffer().view(np.complex64)[: R + (1 if is_real else 0)])
    assert np.array_equal(eng.get_quantized_buffer(), want)
    rng = np.random.default_rng(3)
    levels, ls, rs = [], [], []
    for lv in range(L):  # every level: the default window (websocket.cpp:195-198), edges, an empty slice, random windows
        width = R >> lv
        for (l, r) in [(0, min(width, cfg.waterfall_size)), (0, width), (width - 1, width), (width // 2, width // 2)] + \
                [tuple(sorted(rng.integers(0, width + 1, 2))) for _ in range(3)]:
            levels.append(lv)
            ls.append(int(l))
            rs.append(int(r))
    rows = eng.waterfall_gather(levels, ls, rs)
    for lv, l, r, row in zip(levels, ls, rs, rows):
        off = oracle.level_offset(lv, R)  # websocket.cpp:233: level i starts at sum_{j<i} (R >> j)
        assert np.array_equal(row, want[off + l: off + r]), (lv, l, r)
    # a second, larger call reuses / grows the staging buffers
    rows2 = eng.waterfall_gather([0] * 40, [0] * 40, [R] * 40)
    assert all(np.array_equal(x, want[:R]) for x in rows2)
    eng.close()


@pytest.mark.parametrize("is_real", [False, True])
def test_waterfall_cadence_and_int16_pcm(gpu_required, is_real):
    """Blocks with skip_num = 3 and int16 PCM against blocks with the defaults: PCM values identical, pyramid rows of the
    send frames identical, the other rows untouched; per-frame execute() only refreshes the mirror on send frames."""
    from phantomsdr_b200.backend import OPT_PCM16

    cfg = SpectrumConfig(sps=4_370_000 * (2 if is_real else 1), fft_size=(1 << 18) if is_real else (1 << 17), is_real=is_real)
    n, h = cfg.audio_fft_size, cfg.audio_fft_size // 2
    F, nblocks, nc, skip = 8, 3, 37, 3
    src = SignalSource(cfg, seed=9)
    hops = [hop_as_floats(src.next_hop()).copy() for _ in range(F * nblocks + 1)]
    specs = make_clients(cfg, nc, modes=(USB, LSB, AM, FM), tones=[src.display_bin(t) for t in src.tones])

    def run(cadence, pcm16):
        e = make_engine(cfg)
        e.set_hop_ring(2 * F + 2)
        e.set_batch_frames(F)
        e.set_pipeline(2)
        if pcm16:
            e.set_option(OPT_PCM16, 1)
        e.set_waterfall_cadence(cadence)
        e.clients_create(nc, n, cfg.audio_sps)
        for i, c in enumerate(specs):
            e.client_open(i, c.l, c.mid, c.r, c.mode)
        sets = [dict(halves=[e.malloc(cfg.hop_floats) for _ in range(F)], pcm=e.pinned(4 * F * nc * h, np.int32),
                     pwr=e.pinned(4 * F * nc, np.float32), valid=e.pinned(F * nc, np.uint8),
                     pyr=e.pinned(F * e.pyramid_bytes, np.int8)) for _ in range(2)]
        for st in sets:
            st["pyr"][:] = 77  # sentinel: rows of frames that are not sent must stay untouched
        prime = e.malloc(cfg.hop_floats)
        prime[:] = hops[0]
        e.stream_prime(prime)
        pcm, pyr = [], []
        for k in range(nblocks):
            st = sets[k & 1]
            for f in range(F):
                st["halves"][f][:] = hops[1 + k * F + f]
            e.submit_block(st["halves"], k * F, st["pcm"], st["pwr"], st["valid"], st["pyr"])
            e.wait_block()
            raw = st["pcm"].view(np.int16)[: F * nc * h] if pcm16 else st["pcm"]
            pcm.extend(raw.reshape(F, nc, h).astype(np.int32).copy())
            pyr.extend(st["pyr"].reshape(F, -1).copy())
            st["pyr"][:] = 77
        e.close()
        return pcm, pyr

    ref_pcm, ref_pyr = run(1, False)
    got_pcm, got_pyr = run(skip, True)
    assert any(p.any() for p in ref_pcm[F:]), "AGC never opened: test is vacuous"
    for f in range(F * nblocks):
        assert np.array_equal(got_pcm[f], ref_pcm[f]), f"frame {f}: int16 PCM differs from int32 PCM"
        if f % skip == 0:
            assert np.array_equal(got_pyr[f], ref_pyr[f]), f"frame {f}: pyramid of a send frame differs"
        else:
            assert (got_pyr[f] == 77).all(), f"frame {f}: pyramid row of a skipped frame was written"

    # per-frame reference-shaped calls: the quantized mirror changes only on send frames
    e = make_engine(cfg)
    e.set_waterfall_cadence(skip)
    ring = [e.malloc(cfg.hop_floats) for _ in range(3)]
    ring[0][:] = hops[0]
    last = None
    for f in range(7):
        ring[(f + 1) % 3][:] = hops[f + 1]
        (e.load_real_input if is_real else e.load_complex_input)(ring[f % 3], ring[(f + 1) % 3])
        e.execute()
        q = e.get_quantized_buffer().copy()
        if f % skip == 0:
            assert np.array_equal(q, ref_pyr[f]), f"frame {f}"
        else:
            assert np.array_equal(q, last), f"frame {f}: mirror changed on a frame that is not sent"
        last = q
    e.close()
