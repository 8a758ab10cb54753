"""Multi-GPU plumbing for the waveform path: utterances are independent (even the aspiration-noise stream is
re-derived per utterance, reference src/lib.rs:594), so a batch is sharded by utterance with NO collective on the data
path.  One process per GPU; torch.distributed is used only for (optional) output gathering and timing barriers."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np


def lpt_assign(sample_counts: Sequence[int], world_size: int) -> List[np.ndarray]:
    """Longest-processing-time greedy assignment of utterances to ranks on their exact sample counts
    (grail_cuda_count_samples).  Returns, per rank, the utterance indices it owns (ascending)."""
    counts = np.asarray(sample_counts, dtype=np.int64)
    order = np.argsort(-counts, kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    owner = np.empty(len(counts), dtype=np.int64)
    for u in order:
        r = int(np.argmin(load))
        owner[u] = r
        load[r] += counts[u]
    return [np.flatnonzero(owner == r) for r in range(world_size)]


def shard_batch(elems: np.ndarray, utt_offsets: np.ndarray, voices: np.ndarray, mine: np.ndarray):
    """the sub-batch (elems, utt_offsets, voices) of the utterances in `mine`"""
    offs = np.asarray(utt_offsets, dtype=np.int64)
    parts = [elems[offs[u]:offs[u + 1]] for u in mine]
    lens = [len(p) for p in parts]
    sub = np.concatenate(parts) if parts else elems[:0]
    sub_offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    return np.ascontiguousarray(sub), sub_offs, np.ascontiguousarray(voices[mine])


def gather_outputs(local_out, local_counts: Sequence[int], assignment: List[np.ndarray], all_counts: Sequence[int],
                   group=None):
    """Reassemble the full batch output in utterance order on every rank.

    local_out is this rank's packed output (torch tensor on its device, or numpy on CPU); each rank contributes a
    padded shard to one all_gather (NCCL over NVLink on GPUs, gloo in the CPU tests); <= 4 B/sample, never the limit."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    t = local_out if isinstance(local_out, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local_out))
    all_counts = np.asarray(all_counts, dtype=np.int64)
    shard_sizes = [int(all_counts[a].sum()) for a in assignment]
    pad = max(shard_sizes) if shard_sizes else 0
    buf = torch.zeros(pad, dtype=t.dtype, device=t.device)
    buf[: t.numel()] = t
    gathered = torch.empty(world * pad, dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(gathered, buf, group=group)
    total = int(all_counts.sum())
    full = torch.empty(total, dtype=t.dtype, device=t.device)
    offs = np.concatenate([[0], np.cumsum(all_counts)])
    for r, a in enumerate(assignment):
        pos = r * pad
        for u in a:
            n = int(all_counts[u])
            full[offs[u]: offs[u] + n] = gathered[pos: pos + n]
            pos += n
    return full
