"""Host-side logic that needs no GPU: derived sizes, synthetic generator, client partitioning, and the
world_size-2 exchange step over gloo."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

from phantomsdr_b200 import SpectrumConfig, sizes, USB, LSB, AM, FM
from phantomsdr_b200.parallel import partition_clients, sorted_client_order, block_subband
from phantomsdr_b200.synth import SignalSource, make_clients

ROOT = Path(__file__).resolve().parent.parent


def test_config_table_of_survey():
    c2 = SpectrumConfig(sps=35_000_000, fft_size=1 << 20)
    assert (c2.fft_result_size, c2.audio_fft_size, c2.downsample_levels, c2.skip_num) == (1 << 20, 360, 11, 6)
    assert c2.base_idx == (1 << 19) + 1 and c2.hop_floats == 1 << 20
    assert sizes.pyramid_size(c2.fft_result_size, 11) == 2_096_128
    c3 = SpectrumConfig(sps=70_000_000, fft_size=1 << 21, is_real=True)
    assert (c3.fft_result_size, c3.audio_fft_size, c3.downsample_levels, c3.hop_floats) == (1 << 20, 360, 11, 1 << 20)
    c1 = SpectrumConfig(sps=2_880_000, fft_size=1 << 17)
    assert (c1.audio_fft_size, c1.downsample_levels, c1.skip_num) == (548, 8, 4)
    assert c1.slice_offset(0) == (1 << 16) + 1 and c1.slice_offset((1 << 16) - 1) == 0


def test_synth_is_deterministic_and_below_wrap():
    cfg = SpectrumConfig(sps=35_000_000, fft_size=1 << 16)
    a = SignalSource(cfg, seed=3).next_hop()
    b = SignalSource(cfg, seed=3).next_hop()
    assert np.array_equal(a, b) and a.dtype == np.complex64 and a.size == cfg.hop_samples
    assert max(t.amp for t in SignalSource(cfg, seed=3).tones) < 2 / np.sqrt(cfg.fft_size)
    r = SignalSource(SpectrumConfig(sps=70_000_000, fft_size=1 << 17, is_real=True), seed=3).next_hop()
    assert r.dtype == np.float32


def test_client_table_respects_reference_validation():
    cfg = SpectrumConfig(sps=35_000_000, fft_size=1 << 20)
    cl = make_clients(cfg, 300, modes=(USB, LSB, AM, FM))
    R, n = cfg.fft_result_size, cfg.audio_fft_size
    for c in cl:
        assert 0 <= c.l <= c.r < R and c.r - c.l <= n  # signal.cpp:304-311
        assert c.l <= c.mid <= c.r + 1


def test_partition_is_contiguous_in_multimap_order():
    rng = np.random.default_rng(0)
    clients = [(int(l), int(l + w)) for l, w in zip(rng.integers(0, 10000, 1001), rng.integers(0, 300, 1001))]
    order = sorted_client_order(clients)
    assert [clients[i] for i in order] == sorted(clients)
    for world in (1, 2, 4, 8):
        parts = partition_clients(clients, world)
        assert sum(parts, []) == order  # contiguous blocks, nothing lost or duplicated
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
        for a, b in zip(parts, parts[1:]):
            if a and b:
                assert clients[a[-1]] <= clients[b[0]]
        lo, hi = block_subband(clients, parts[-1])
        assert lo == clients[parts[-1][0]][0] and hi >= clients[parts[-1][-1]][1]


def test_two_rank_exchange_over_gloo():
    """world_size 2, gloo, CPU: rank 0 'ingests' (oracle spectrum), one broadcast, each rank demodulates its
    block; the union must equal the single-process result and every rank's frame must hash equal."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611", PYTHONPATH=str(ROOT))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29611", str(ROOT / "tests" / "gloo_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "GLOO_EXCHANGE_OK" in out.stdout
