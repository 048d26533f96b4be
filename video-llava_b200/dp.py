"""Data-parallel plumbing for the hot path: clips are independent end to end, so
the only multi-GPU logic is (1) which clips a rank owns and (2) ONE all_gather of the generated
token ids. Works with any torch.distributed backend (nccl on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import torch


def shard_clips(n_clips: int, rank: int, world: int) -> list[int]:
    """Round-robin: clip i belongs to rank i % world. Every clip is owned exactly once."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_clips, world))


def gather_tokens(local: torch.Tensor, n_clips: int, rank: int, world: int, dist=None) -> torch.Tensor:
    """local: [n_local, n_new] int32 token ids of this rank's clips (in shard_clips order).
    Returns [n_clips, n_new] in clip order on every rank. Ranks with fewer clips are padded to the
    common maximum for the single all_gather and the padding is dropped afterwards."""
    if world == 1:
        return local
    per = (n_clips + world - 1) // world
    n_new = local.shape[1]
    buf = torch.full((per, n_new), -1, dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty(world * per, n_new, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, buf)
    out = out.view(world, per, n_new)
    res = torch.empty(n_clips, n_new, dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = shard_clips(n_clips, r, world)
        res[idx] = out[r, : len(idx)]
    return res
