"""Multi-GPU sharding of the client path: one process per GPU; the ingest rank runs
the forward FFT, ONE collective delivers each spectrum batch to every rank (NCCL broadcast over
NVLink; gloo in the CPU tests), every rank demodulates a contiguous block of the (l, r)-sorted
client list. There is no reference code for this (the reference is single-process); the client
order is the reference's multimap order (src/spectrumserver.h:164-165).

torch.distributed is plumbing only: the tensors broadcast are zero-copy views of engine-owned device
memory (backend._DevArray) or, in the CPU tests, plain host tensors.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple


def sorted_client_order(clients: Sequence[Tuple[int, int]]) -> List[int]:
    """Indices of clients in std::multimap<pair<l, r>> iteration order (stable for equal keys)."""
    return sorted(range(len(clients)), key=lambda i: (clients[i][0], clients[i][1]))


def partition_clients(clients: Sequence[Tuple[int, int]], world: int) -> List[List[int]]:
    """Contiguous, equal-count (+-1) blocks of the (l, r)-sorted client list, one block per rank.
    Sticky by construction: a rank's block only changes when clients connect / disconnect."""
    order = sorted_client_order(clients)
    n = len(order)
    base, extra = divmod(n, world)
    out, pos = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        out.append(order[pos:pos + cnt])
        pos += cnt
    return out


def block_subband(clients: Sequence[Tuple[int, int]], block: Sequence[int]) -> Tuple[int, int]:
    """[min l, max r) of a rank's block: the only display bins that rank ever reads (ablation (ii) of
    SURVEY 8e: send/recv just this sub-band instead of broadcasting the full frame)."""
    if not block:
        return (0, 0)
    return (min(clients[i][0] for i in block), max(clients[i][1] for i in block))


@dataclass
class SpectrumExchange:
    """The per-batch exchange step. `src` is the ingest rank."""

    world: int
    rank: int
    src: int = 0

    def broadcast(self, spectrum_tensor):
        """Deliver the ingest rank's spectrum batch to every rank, in place."""
        if self.world == 1:
            return spectrum_tensor
        import torch.distributed as dist

        dist.broadcast(spectrum_tensor, src=self.src)
        return spectrum_tensor

    def checksum_agrees(self, spectrum_tensor) -> bool:
        """Test aid: every rank's frame must be bit-identical to the ingest rank's."""
        if self.world == 1:
            return True
        import torch
        import torch.distributed as dist

        flat = spectrum_tensor.detach().reshape(-1).view(torch.int32).to(torch.int64)
        local = torch.stack([flat.sum(), (flat * (torch.arange(flat.numel(), device=flat.device) % 8191 + 1)).sum()])
        ref = local.clone()
        dist.broadcast(ref, src=self.src)
        ok = torch.tensor([int(bool((ref == local).all()))], device=flat.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        return bool(ok.item())
