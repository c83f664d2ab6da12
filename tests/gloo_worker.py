"""Worker of tests/test_host_logic.py::test_two_rank_exchange_over_gloo (launched by torchrun, gloo, CPU)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import oracle  # noqa: E402
from phantomsdr_b200 import SpectrumConfig, USB, LSB, AM  # noqa: E402
from phantomsdr_b200.parallel import SpectrumExchange, partition_clients  # noqa: E402
from phantomsdr_b200.synth import SignalSource, make_clients  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    cfg = SpectrumConfig(sps=546_000, fft_size=1 << 14)  # n = 360 at a small FFT (oracle only, no GPU here)
    n, R = cfg.audio_fft_size, cfg.fft_result_size
    specs = make_clients(cfg, 21, modes=(AM, USB, LSB))
    parts = partition_clients([(c.l, c.r) for c in specs], world)
    ex = SpectrumExchange(world, rank)
    mine = {i: oracle.OracleClient(False, n, cfg.audio_sps, R) for i in parts[rank]}
    everyone = {i: oracle.OracleClient(False, n, cfg.audio_sps, R) for i in range(len(specs))} if rank == 0 else {}
    for table in (mine, everyone):
        for i, c in table.items():
            c.set_audio_range(specs[i].l, specs[i].mid, specs[i].r)
            c.set_audio_demodulation(specs[i].mode)
    orc = oracle.OracleFFT(cfg.fft_size, cfg.downsample_levels, 0)
    orc.set_output_additional_size(n)
    orc.plan_c2c()
    src = SignalSource(cfg, seed=5)
    hops = [src.next_hop() for _ in range(5)]
    frame_t = torch.zeros(2 * (R + n), dtype=torch.float32)
    for f in range(4):
        if rank == 0:  # ingest rank
            orc.load_complex_input(hops[f], hops[f + 1])
            orc.execute()
            orc.wrap_copy(n)
            frame_t.copy_(torch.from_numpy(orc.outbuf.copy()))
        ex.broadcast(frame_t)
        assert ex.checksum_agrees(frame_t)
        spec = frame_t.numpy().view(np.complex64)
        local = {i: c.send_audio(spec, cfg.fft_size, f)[1] for i, c in mine.items()}
        gathered = [None] * world
        dist.all_gather_object(gathered, local)
        if rank == 0:
            merged = {}
            for g in gathered:
                merged.update(g)
            assert sorted(merged) == list(range(len(specs)))
            for i, c in everyone.items():
                want = c.send_audio(spec, cfg.fft_size, f)[1]
                assert np.array_equal(merged[i], want), (f, i)
    dist.barrier()
    if rank == 0:
        print("GLOO_EXCHANGE_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
