"""Utterance-batch sharding across the GPUs of one box.

The path shards by utterance with no exchange step inside it (every op is per utterance or per
frame), so multi-GPU inference is: partition -> each rank runs its shard on its own engine ->
one gather of token ids / scores.  One process per GPU (``torch.distributed``: NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests); weights are replicated.  The reference has no
multi-GPU inference at all (only DDP training, masr/trainer.py:525,541).
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def partition(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Length-balanced split: sort by length (descending, stable) and deal to the ranks in snake order so
    both sum(T) and sum(T^2) balance; equal lengths degenerate to a contiguous-looking round robin."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    shards: List[List[int]] = [[] for _ in range(world)]
    for pos, idx in enumerate(order):
        rnd, k = divmod(pos, world)
        r = k if rnd % 2 == 0 else world - 1 - k
        shards[r].append(idx)
    return shards


def pack_results(tokens: Sequence[Sequence[int]], scores: Sequence[float], indices: Sequence[int], rows: int, width: int,
                 device) -> Tuple[torch.Tensor, torch.Tensor]:
    """Fixed-shape payload for all_gather: int32 [rows, width+2] = (global index, n tokens, tokens...) and
    float64 [rows] scores; unused rows carry index -1."""
    buf = np.full((rows, width + 2), -1, np.int32)
    sc = np.zeros(rows, np.float64)
    for r, (gi, tk, s) in enumerate(zip(indices, tokens, scores)):
        buf[r, 0] = gi
        buf[r, 1] = len(tk)
        buf[r, 2:2 + len(tk)] = tk
        sc[r] = s
    return torch.from_numpy(buf).to(device), torch.from_numpy(sc).to(device)


def unpack_results(gathered_i: Sequence[torch.Tensor], gathered_s: Sequence[torch.Tensor], total: int):
    tokens: List[List[int]] = [[] for _ in range(total)]
    scores: List[float] = [0.0] * total
    for bi, bs in zip(gathered_i, gathered_s):
        bi, bs = bi.cpu().numpy(), bs.cpu().numpy()
        for r in range(bi.shape[0]):
            gi = int(bi[r, 0])
            if gi < 0:
                continue
            n = int(bi[r, 1])
            tokens[gi] = bi[r, 2:2 + n].tolist()
            scores[gi] = float(bs[r])
    return tokens, scores


def sharded_transcribe(waves: Sequence[np.ndarray], run_local: Callable[[List[np.ndarray]], Tuple[list, list]],
                       max_tokens: int, device="cpu", group=None):
    """SPMD entry point: every rank calls it with the same list; each transcribes its shard with
    ``run_local(list of waveforms) -> (tokens, scores)`` and all ranks return the full, ordered result."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    shards = partition([len(w) for w in waves], world)
    mine = shards[rank]
    tokens, scores = run_local([waves[i] for i in mine]) if mine else ([], [])
    if world == 1:
        out_t: List[List[int]] = [[] for _ in waves]
        out_s = [0.0] * len(waves)
        for i, t, s in zip(mine, tokens, scores):
            out_t[i], out_s[i] = list(t), s
        return out_t, out_s
    rows = max(len(s) for s in shards)
    bi, bs = pack_results(tokens, scores, mine, rows, max_tokens, device)
    gi = [torch.empty_like(bi) for _ in range(world)]
    gs = [torch.empty_like(bs) for _ in range(world)]
    dist.all_gather(gi, bi, group=group)
    dist.all_gather(gs, bs, group=group)
    return unpack_results(gi, gs, len(waves))
