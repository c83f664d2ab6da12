"""GPU parity: batched per-client demodulation through the C ABI vs the CPU oracle's restatement of
AudioClient::send_audio (src/signal.cpp:102-298).

Both sides consume the SAME spectrum (the engine's host mirror, wrap tail included), so the
comparison isolates the client chain:
  * slice offsets / placement / parity flip / overlap-add / demod: audio before DC removal within
    1e-5 * max|y| per frame (north_star tolerance; FM compared as a wrapped angle);
  * DC blocker + AGC + int16: BIT-EXACT when the oracle tails are fed the engine's own pre-DC audio
    (strictly sequential float recurrences restated op-for-op), and within 1 LSB against the full
    oracle chain (measured: 1 LSB on fewer than 0.1 % of the samples);
  * pwr within 1e-5 relative.
"""
import numpy as np
import pytest

import oracle
from phantomsdr_b200 import SpectrumConfig, USB, LSB, AM, FM
from phantomsdr_b200.synth import SignalSource, make_clients, ClientSpec
from helpers import make_engine, make_oracle_fft, hop_as_floats

pytestmark = pytest.mark.gpu


class OracleTail:
    def __init__(self, cfg):
        self.dc = oracle.OracleDC(cfg.audio_sps // 750 * 2)
        self.agc = oracle.OracleAGC(0.2, 50.0, 300.0, 200.0, float(cfg.audio_sps))

    def run(self, pre):
        y = self.agc.process(self.dc.remove(pre))
        out = np.zeros(y.size, np.int32)
        oracle.lib().orc_float_to_int16(y, out, 16384.0, y.size)
        return out


def run_case(cfg, clients, nframes, mode_changes=None, seed=7):
    src = SignalSource(cfg, seed=seed)
    eng = make_engine(cfg)
    n = cfg.audio_fft_size
    h = n // 2
    nc = len(clients)
    eng.clients_create(nc + 3, n, cfg.audio_sps)  # a few spare, closed slots
    orcs, tails = [], []
    for i, c in enumerate(clients):
        eng.client_open(i, c.l, c.mid, c.r, c.mode)
        o = oracle.OracleClient(cfg.is_real, n, cfg.audio_sps, cfg.fft_result_size)
        assert o.on_window_message(c.l, c.mid, c.r)
        o.set_audio_demodulation(c.mode)
        orcs.append(o)
        tails.append(OracleTail(cfg))
    ring = [eng.malloc(cfg.hop_floats) for _ in range(3)]
    ring[0][:] = hop_as_floats(src.next_hop())
    ring[1][:] = hop_as_floats(src.next_hop())
    idx = 0
    stats = {"pre_rel": 0.0, "pcm_max": 0, "pcm_diff_frac": 0.0}
    total = diff = 0
    for frame in range(nframes):
        if mode_changes and frame in mode_changes:
            for (ci, mode) in mode_changes[frame]:
                eng.client_set_demodulation(ci, mode)
                orcs[ci].on_demodulation_message(mode)
                tails[ci].agc.reset()
                clients[ci] = ClientSpec(clients[ci].l, clients[ci].mid, clients[ci].r, mode)
        a1, a2 = ring[idx], ring[(idx + 1) % 3]
        (eng.load_real_input if cfg.is_real else eng.load_complex_input)(a1, a2)
        ring[(idx + 2) % 3][:] = hop_as_floats(src.next_hop())
        idx = (idx + 1) % 3
        eng.execute()
        spec = eng.get_output_buffer().view(np.complex64).copy()
        pcm, pwr, valid = eng.clients_execute(frame)
        pre = eng.clients_read_pre_dc()
        assert not valid[nc:].any(), "closed slots must read back invalid"
        for i, (c, o) in enumerate(zip(clients, orcs)):
            ok, pcm_ref, pwr_ref, pre_ref = o.send_audio(spec, cfg.fft_size, frame)
            assert ok and valid[i] == 1
            scale = max(float(np.abs(pre_ref).max()), 1e-30)
            if c.mode == FM:
                d = np.angle(np.exp(1j * (pre[i].astype(np.float64) - pre_ref)))
                assert (np.abs(d) > 1e-3).mean() <= 0.01, f"client {i} FM: angle error"
            else:
                err = float(np.abs(pre[i] - pre_ref).max()) / scale
                stats["pre_rel"] = max(stats["pre_rel"], err)
                assert err <= 1e-5, f"frame {frame} client {i} mode {c.mode}: pre-DC audio rel err {err:.2e}"
            assert abs(pwr[i] - pwr_ref) <= 1e-5 * max(pwr_ref, 1e-30), f"client {i}: pwr {pwr[i]} vs {pwr_ref}"
            # sequential tails: bit-exact on identical input
            exact = tails[i].run(pre[i])
            assert np.array_equal(pcm[i], exact), f"frame {frame} client {i}: DC/AGC/int16 tail not bit-exact"
            dd = np.abs(pcm[i] - pcm_ref)
            if c.mode != FM:
                stats["pcm_max"] = max(stats["pcm_max"], int(dd.max()))
                total += dd.size
                diff += int((dd != 0).sum())
            else:
                # FM through the whole oracle chain too: a discriminator sample that lands on the other side of +-pi is a
                # 2 pi step into the DC blocker and the AGC, so the bar is statistical - at most 1 % of the samples off by
                # more than 2 LSB
                stats["fm_total"] = stats.get("fm_total", 0) + dd.size
                stats["fm_off"] = stats.get("fm_off", 0) + int((dd > 2).sum())
    stats["pcm_diff_frac"] = diff / max(total, 1)
    stats["fm_off_frac"] = stats.get("fm_off", 0) / max(stats.get("fm_total", 0), 1)
    eng.close()
    return stats


def cfg_for_n(fft_size, n_target, is_real=False, audio_sps=12000):
    # pick sps so that audio_fft_size == n_target (src/spectrumserver.cpp:151)
    sps = int(audio_sps * fft_size / (n_target - 1.5))
    cfg = SpectrumConfig(sps=sps, fft_size=fft_size, is_real=is_real, audio_sps=audio_sps)
    assert cfg.audio_fft_size == n_target, cfg.audio_fft_size
    return cfg


@pytest.mark.parametrize("n_target,fft_size,is_real", [
    (360, 1 << 17, False),   # cfg 2's audio size: 2^3 * 3^2 * 5
    (548, 1 << 17, False),   # cfg 1 at 2.88 MSPS: 4 * 137 (generic prime radix)
    (492, 1 << 16, False),   # cfg 1 at 3.2 MSPS: 4 * 3 * 41
    (360, 1 << 18, True),    # real input: base_idx 0, opposite parity rule
])
def test_clients_all_modes(gpu_required, n_target, fft_size, is_real):
    cfg = cfg_for_n(fft_size, n_target, is_real)
    src = SignalSource(cfg, seed=7)
    tones = [src.display_bin(t) for t in src.tones]
    clients = make_clients(cfg, 24, modes=(USB, LSB, AM, FM), tones=tones, on_tone_fraction=0.7)
    stats = run_case(cfg, clients, nframes=18)
    # measured on B200: at most 1 LSB, on fewer than 0.1 % of the samples; no FM sample off by more than 2 LSB
    assert stats["pcm_max"] <= 1 and stats["pcm_diff_frac"] <= 0.005, stats
    assert stats["fm_off_frac"] <= 0.002, stats


def test_clients_edges_and_mode_switch(gpu_required):
    """Edge slices: l near 0 (IQ wrap through the tail), r-l = n (max), r = l (empty), odd/even
    floor(mid) for both frame parities, slice not containing mid, demodulation switches with AGC
    reset (src/signal.cpp:316-328)."""
    cfg = cfg_for_n(1 << 17, 360)
    R, n = cfg.fft_result_size, cfg.audio_fft_size
    half_wrap = R - (cfg.fft_size // 2 + 1)  # display l whose slice starts at FFT bin R-1 ... wraps into the tail
    clients = [
        ClientSpec(half_wrap - 40, half_wrap + 10.25, half_wrap - 40 + n, USB),
        ClientSpec(half_wrap - 3, half_wrap + 100.5, half_wrap + 200, AM),
        ClientSpec(0, 0.0, 90, USB),
        ClientSpec(0, 45.5, 90, LSB),
        ClientSpec(1000, 1001.0, 1000, USB),          # empty slice
        ClientSpec(2000, 2150.75, 2300, FM),
        ClientSpec(3001, 3001.5, 3091, USB),          # odd floor(mid)
        ClientSpec(3000, 3000.5, 3090, USB),          # even floor(mid)
        ClientSpec(5000, 5400.0, 5100, USB),          # mid outside the slice
        ClientSpec(R - 1 - n, R - 200.0, R - 1, LSB), # top edge
    ]
    changes = {5: [(0, LSB), (5, AM)], 9: [(0, AM), (1, FM), (5, USB)], 12: [(0, USB)]}
    stats = run_case(cfg, clients, nframes=16, mode_changes=changes)
    assert stats["pcm_max"] <= 1 and stats["fm_off_frac"] <= 0.002, stats   # (measured: 0 and 0)


def test_wbfm_width_clients(gpu_required):
    """WBFM-width clients (default window +-96 kHz at 192 kHz audio, src/spectrumserver.cpp:137-140): a 10 488-point
    audio FFT (2^3 * 3 * 19 * 23: generic prime radices), slices of ~10 000 bins and a 512-sample DC blocker - the sizes
    at which the engine falls back from the warp-per-client demodulation and the pipelined tails to its general
    kernels. Same bars as the narrow-band cases."""
    cfg = SpectrumConfig(sps=2_400_000, fft_size=1 << 17, audio_sps=192000)
    n = cfg.audio_fft_size
    assert n == 10488
    R = cfg.fft_result_size
    half = int(96000 / (cfg.sps / cfg.fft_size))  # +-96 kHz in bins
    mids = [R // 4 + 0.25, R // 2 + 1000.5, 3 * R // 4 + 7.0]
    clients = [ClientSpec(int(m) - half, m, int(m) + half, FM) for m in mids]
    clients.append(ClientSpec(int(mids[0]) - half // 2, mids[0] + 3.0, int(mids[0]) + half // 2, AM))
    stats = run_case(cfg, clients, nframes=6)
    assert stats["pcm_max"] <= 1 and stats["fm_off_frac"] <= 0.002, stats   # (measured: 0 and 0)


def test_clients_create_can_be_repeated_after_a_failure(gpu_required):
    """A client table that fails half way (here: an audio FFT too long for the demodulation kernels, found after the
    state arrays have been allocated) leaves nothing behind: the call can be repeated with a supported size and the
    clients then behave as in any other test."""
    from phantomsdr_b200.backend import B200Error

    cfg = cfg_for_n(1 << 17, 360)
    eng = make_engine(cfg)
    with pytest.raises(B200Error):
        eng.clients_create(8, 16384, cfg.audio_sps)   # rejected before anything is allocated (additional size < n)
    eng.close()

    # the failure that happens AFTER allocation needs additional >= n: a fresh engine planned with a large tail
    from phantomsdr_b200.backend import B200FFT
    e2 = B200FFT(cfg.fft_size, 1, cfg.downsample_levels, cfg.brightness_offset, 0)
    e2.set_output_additional_size(16384)
    e2.plan_c2c()
    with pytest.raises(B200Error, match="too large"):
        e2.clients_create(8, 16384, 192000)   # (192 kHz audio: the AGC look-ahead check passes, the size check does not)
    e2.clients_create(8, 360, cfg.audio_sps)        # no "clients already created", no leak of the first attempt
    e2.client_open(0, 1000, 1001.0, 1090, USB)
    ring = [e2.malloc(cfg.hop_floats) for _ in range(2)]
    src = SignalSource(cfg, seed=3)
    ring[0][:] = hop_as_floats(src.next_hop())
    ring[1][:] = hop_as_floats(src.next_hop())
    e2.load_complex_input(ring[0], ring[1])
    e2.execute()
    pcm, pwr, valid = e2.clients_execute(0)
    assert valid[0] == 1 and not valid[1:].any()
    e2.close()
