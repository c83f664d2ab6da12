"""Worker of tests/test_gpu_multi.py (torchrun, NCCL, one rank per GPU): the spectrum exchange step.
Every rank computes a LOCAL reference (its own forward FFT on the same input + its client block) and then the same
client block on the spectrum delivered by the ingest rank, once by NCCL broadcast and once by the fused peer-store
scatter of FFT pass 2 (CUDA IPC + stream-ordered flags). PCM must be bit-identical in all three."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from phantomsdr_b200 import SpectrumConfig, USB, LSB, AM, FM  # noqa: E402
from phantomsdr_b200.backend import B200FFT, OPT_PEER_STORES  # noqa: E402
from phantomsdr_b200.parallel import partition_clients, SpectrumExchange  # noqa: E402
from phantomsdr_b200.synth import SignalSource, make_clients  # noqa: E402


def make(cfg, dev, F, nhops):
    e = B200FFT(cfg.fft_size, 1, cfg.downsample_levels, 0, dev)
    e.set_output_additional_size(cfg.audio_fft_size)
    e.plan_r2c() if cfg.is_real else e.plan_c2c()
    e.set_hop_ring(nhops)
    e.set_batch_frames(F)
    e.set_pipeline(2)
    return e


def CASES():
    return (SpectrumConfig(sps=35_000_000, fft_size=1 << 20), SpectrumConfig(sps=70_000_000, fft_size=1 << 21, is_real=True))


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    # c2c (BASELINE cfg 2 shape) and r2c (cfg 3 shape: the Hermitian-split kernel writes the spectrum and the peer copies)
    for cfg in CASES():
        exchange_case(cfg, local, dev, rank, world)
    dist.barrier()
    if rank == 0:
        print("MGPU_EXCHANGE_OK")
    dist.destroy_process_group()


def exchange_case(cfg, local, dev, rank, world):
    # (failures are collected and raised at the end: every rank walks through the same collectives whatever it finds)
    problems = []
    n, h, F, nb = cfg.audio_fft_size, cfg.audio_fft_size // 2, 4, 3
    nc = 24
    everyone = make_clients(cfg, nc * world, seed=77, modes=(USB, LSB, AM, FM))
    parts = partition_clients([(c.l, c.r) for c in everyone], world)
    mine = [everyone[i] for i in parts[rank]]
    src = SignalSource(cfg, seed=5, ntones=8)
    hops = np.stack([src.next_hop().view(np.float32) for _ in range(F * nb + 1)])

    def run(mode):
        e = make(cfg, local, F, F * nb + 1)
        e.clients_create(nc, n, cfg.audio_sps)
        for i, c in enumerate(mine):
            e.client_open(i, c.l, c.mid, c.r, c.mode)
        stream = torch.cuda.Stream(device=dev)
        torch.cuda.set_stream(stream)
        e.set_stream(stream.cuda_stream)
        ring = torch.as_tensor(e.device_hop_ring(F * nb + 1), device=dev)
        if mode == "local" or rank == 0:
            ring.copy_(torch.from_numpy(hops))
        banks = []
        for b in range(2):
            e.select_bank(b)
            banks.append(torch.as_tensor(e.device_spectrum(F), device=dev))
        peer_ready, r0_consumed, my_ready, my_consumed = [], [], None, None
        remote, my_range = None, None

        def block_ranges(g):  # the two half-open bin ranges (below / above the wrap point of the display axis) rank g needs
            blk = [everyone[i] for i in parts[g]]
            iv = sorted((cfg.slice_offset(c.l), cfg.slice_offset(c.l) + c.r - c.l) for c in blk)
            half = cfg.fft_result_size // 2
            low = [x for x in iv if x[0] < half]
            high = [x for x in iv if x[0] >= half]
            return [(min(a for a, _ in part), max(b for _, b in part)) if part else (0, 0) for part in (low, high)]

        if mode in ("scatter", "scatter-dma", "scatter-pull"):
            flags = e.flag_buffer
            table = [None] * world
            dist.all_gather_object(table, {"spec": e.ipc_export(e.spectrum_base), "flags": e.ipc_export(flags)})
            if rank == 0:
                ptrs = []
                for g in range(1, world):
                    if mode != "scatter-pull":
                        ptrs.append(e.ipc_open(table[g]["spec"]) + e.spectrum_offset)
                    peer_ready.append(e.ipc_open(table[g]["flags"]))
                    r0_consumed.append(flags + 8 * g)
                if mode != "scatter-pull":
                    e.set_peer_spectra(ptrs)
                    for g in range(1, world):
                        r = block_ranges(g)
                        e.set_peer_ranges(g - 1, r[0][0], r[0][1], r[1][0], r[1][1])
                if mode == "scatter-dma":
                    e.set_option(OPT_PEER_STORES, 0)
            else:
                my_ready = flags
                my_consumed = e.ipc_open(table[0]["flags"]) + 8 * rank
                if mode == "scatter-pull":  # this rank's copy engine fetches its sub-band from the ingest rank's bank
                    remote = e.ipc_open(table[0]["spec"]) + e.spectrum_offset
                    my_range = block_ranges(rank)
            dist.barrier()
        ex = SpectrumExchange(world, rank)
        out = []
        for k in range(nb):
            bank = k % 2
            e.select_bank(bank)
            seq = k + 1
            if mode == "local":
                e.execute_device(k * F, F)
            elif mode == "broadcast":
                if rank == 0:
                    e.execute_device(k * F, F)
                else:
                    e.bank_acquire()
                ex.broadcast(banks[bank])
                if not ex.checksum_agrees(banks[bank]):
                    problems.append(f"rank {rank} batch {k}: broadcast checksum differs")
            elif mode == "scatter-dma" and rank == 0:
                e.execute_device(k * F, F)
                if seq > 2:
                    e.enqueue_wait(2, r0_consumed, seq - 2)
                e.push_peers(F)
                e.enqueue_signal(2, peer_ready, seq)
            else:
                if rank == 0:
                    if seq > 2:
                        e.enqueue_wait(False, r0_consumed, seq - 2)
                    e.execute_device(k * F, F)
                    e.enqueue_signal(False, peer_ready, seq)
                else:
                    e.enqueue_wait(True, [my_ready], seq)
                    if mode == "scatter-pull":
                        e.pull_spectrum(remote, F, my_range[0][0], my_range[0][1], my_range[1][0], my_range[1][1])
                        e.enqueue_signal(True, [my_consumed], seq)  # the ingest rank may reuse its bank
            e.clients_execute_device(k * F, F)
            if mode in ("scatter", "scatter-dma") and rank != 0:
                e.enqueue_signal(True, [my_consumed], seq)
            for f in range(F):
                pcm, pwr, valid = e.clients_fetch(f)
                if not valid[:nc].all():
                    problems.append(f"rank {rank} mode {mode} batch {k} frame {f}: invalid client frames")
                out.append(pcm.copy())
        e.sync()
        if e.flag_error != 0:
            problems.append(f"rank {rank} mode {mode}: a peer flag wait timed out")
        if mode in ("scatter", "scatter-dma", "scatter-pull"):
            dist.barrier()
        e.close()
        return out

    ref = run("local")
    for mode in ("broadcast", "scatter", "scatter-dma", "scatter-pull"):
        got = run(mode)
        for f, (a, b) in enumerate(zip(ref, got)):
            if not np.array_equal(a, b):
                problems.append(f"rank {rank} {'r2c' if cfg.is_real else 'c2c'} mode {mode} frame {f}: PCM differs")
                break
    dist.barrier()
    assert not problems, problems[:4]


if __name__ == "__main__":
    main()
