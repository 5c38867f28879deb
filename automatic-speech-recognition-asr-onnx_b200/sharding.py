"""Utterance-level data parallelism: one process + one engine per GPU, weights replicated,
no collective inside the model; ONE all-gather of the padded token matrix at the very end
(<= 57 KB for 32 x 448 int32 -- pure latency over NVLink/NVSwitch).

The reference has no multi-device path at all: its driver runs clips
strictly one after another (Whisper/Inference_Whisper_ONNX.py:722-768).  Clips and sliding
windows share no state, so they shard freely.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_indices(lengths: Sequence[int], world: int, rank: int) -> List[int]:
    """Sort clips by length (longest first) and deal them round-robin so every rank gets
    clips of similar duration; returns the global indices owned by `rank` (ascending)."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    return sorted(order[rank::world])


def group_by_length(indices: Sequence[int], lengths: Sequence[int]) -> List[Tuple[int, List[int]]]:
    """Batches of equal sample count (the uniform `encode` / `transcribe` calls); mixed lengths go through
    `ragged_batches` and the `lens=` arguments instead."""
    groups = {}
    for i in indices:
        groups.setdefault(int(lengths[i]), []).append(i)
    return sorted(groups.items())


def ragged_batches(indices: Sequence[int], lengths: Sequence[int], max_batch: int) -> List[List[int]]:
    """Batches for `WhisperEngine.transcribe(..., lens=)`: the engine gives every clip of a ragged batch its single-clip
    semantics (own reflect pad, mel maximum, attention key range), so clips of any lengths may share a batch; neighbours in
    length order are put together because the batch runs at its longest clip's row count."""
    order = sorted(indices, key=lambda i: (-int(lengths[i]), i))
    return [order[k:k + max_batch] for k in range(0, len(order), max_batch)]


def gather_tokens(local_tokens: Sequence[Sequence[int]], local_indices: Sequence[int], n_total: int, max_len: int,
                  device: torch.device | str = "cpu", group=None) -> List[List[int]]:
    """All ranks end up with every clip's tokens, in global clip order."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    per_rank = (n_total + world - 1) // world
    buf = torch.full((per_rank, max_len + 2), -1, dtype=torch.int32)
    for row, (idx, toks) in enumerate(zip(local_indices, local_tokens)):
        toks = list(toks)[:max_len]
        buf[row, 0] = idx
        buf[row, 1] = len(toks)
        if toks:
            buf[row, 2:2 + len(toks)] = torch.tensor(toks, dtype=torch.int32)
    if world == 1:
        gathered = buf.unsqueeze(0)
    else:
        buf = buf.to(device)
        out = torch.empty((world, per_rank, max_len + 2), dtype=torch.int32, device=buf.device)
        dist.all_gather_into_tensor(out.view(world * per_rank, max_len + 2), buf, group=group)
        gathered = out.cpu()
    result: List[List[int]] = [[] for _ in range(n_total)]
    g = gathered.reshape(-1, max_len + 2).numpy()
    for row in g:
        if row[0] >= 0:
            result[int(row[0])] = row[2:2 + int(row[1])].tolist()
    return result
