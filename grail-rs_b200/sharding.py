"""Multi-GPU plumbing for the waveform path: utterances are independent (even the aspiration-noise stream is
re-derived per utterance, reference src/lib.rs:594), so a batch is sharded by utterance with NO collective on the data
path.  One process per GPU; torch.distributed is used only for (optional) output gathering and timing barriers."""
from __future__ import annotations

import heapq
from typing import List, Optional, Sequence

import numpy as np


def lpt_assign(sample_counts: Sequence[int], world_size: int) -> List[np.ndarray]:
    """Longest-processing-time greedy assignment of utterances to ranks on their exact sample counts
    (grail_cuda_count_samples).  Returns, per rank, the utterance indices it owns (ascending)."""
    counts = np.asarray(sample_counts, dtype=np.int64)
    order = np.argsort(-counts, kind="stable")
    owner = np.empty(len(counts), dtype=np.int64)
    heap = [(0, r) for r in range(world_size)]          # (load, rank): ties go to the lowest rank, like argmin
    for u in order.tolist():
        load, r = heapq.heappop(heap)
        owner[u] = r
        heapq.heappush(heap, (load + int(counts[u]), r))
    return [np.flatnonzero(owner == r) for r in range(world_size)]


def shard_batch(elems: np.ndarray, utt_offsets: np.ndarray, voices: np.ndarray, mine: np.ndarray):
    """the sub-batch (elems, utt_offsets, voices) of the utterances in `mine`"""
    offs = np.asarray(utt_offsets, dtype=np.int64)
    mine = np.asarray(mine, dtype=np.int64)
    lens = offs[mine + 1] - offs[mine]
    sub_offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    # element index of every phoneme of the sub-batch, without a per-utterance loop
    idx = np.repeat(offs[mine] - sub_offs[:-1], lens) + np.arange(int(sub_offs[-1]), dtype=np.int64)
    return np.ascontiguousarray(elems[idx]), sub_offs.astype(np.uint32), np.ascontiguousarray(voices[mine])


def gather_tables(assignment: List[np.ndarray], all_counts: Sequence[int], pad: int):
    """(dst_off, src_off, len) of every utterance: where it sits in the all-gathered buffer (rank r's packed shard starts
    at r * pad) and where it goes in the batch-ordered output"""
    all_counts = np.asarray(all_counts, dtype=np.int64)
    dst_all = np.concatenate([[0], np.cumsum(all_counts)])[:-1]
    dst, src, ln = [], [], []
    for r, a in enumerate(assignment):
        a = np.asarray(a, dtype=np.int64)
        n = all_counts[a]
        dst.append(dst_all[a])
        src.append(r * pad + np.concatenate([[0], np.cumsum(n)])[:-1])
        ln.append(n)
    cat = lambda x: np.ascontiguousarray(np.concatenate(x) if x else np.zeros(0, np.int64), dtype=np.uint64)   # noqa: E731
    return cat(dst), cat(src), cat(ln)


def _copy_segments_torch(full, gathered, dst, src, ln, chunk: int = 1 << 24):
    """index-copy of whole groups of segments at a time (CPU / gloo path of the tests; no per-utterance loop)"""
    import torch
    dst, src, ln = (torch.from_numpy(x.astype(np.int64)) for x in (dst, src, ln))
    cum = torch.cumsum(ln, 0)
    k0 = 0
    while k0 < len(ln):
        base = int(cum[k0 - 1]) if k0 else 0
        k1 = int(torch.searchsorted(cum, torch.tensor(base + chunk), right=True))
        k1 = max(k1, k0 + 1)
        l = ln[k0:k1]
        seg = torch.repeat_interleave(torch.arange(k1 - k0), l)
        within = torch.arange(int(l.sum())) - torch.repeat_interleave(torch.cumsum(l, 0) - l, l)
        full[(dst[k0:k1][seg] + within).to(full.device)] = gathered[(src[k0:k1][seg] + within).to(full.device)]
        k0 = k1
    return full


def gather_outputs(local_out, local_counts: Sequence[int], assignment: List[np.ndarray], all_counts: Sequence[int],
                   group=None, ctx=None):
    """Reassemble the full batch output in utterance order on every rank.

    local_out is this rank's packed output (torch tensor on its device, or numpy on CPU); each rank contributes a
    padded shard to ONE all_gather (NCCL over NVLink on GPUs, gloo in the CPU tests; <= 4 B/sample, never the limit),
    then one segment-copy kernel of the library (grail_cuda_copy_segments, `ctx` = this rank's grail Context) puts the
    utterances in batch order -- on CPU tensors a chunked index copy does the same."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    t = local_out if isinstance(local_out, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local_out))
    all_counts = np.asarray(all_counts, dtype=np.int64)
    shard_sizes = [int(all_counts[a].sum()) for a in assignment]
    assert t.numel() == shard_sizes[dist.get_rank(group)] == int(np.asarray(local_counts, dtype=np.int64).sum())
    pad = max(shard_sizes) if shard_sizes else 0
    pad = (pad + 3) & ~3                               # shards start 16-byte aligned in the gathered buffer
    buf = torch.zeros(pad, dtype=t.dtype, device=t.device)
    buf[: t.numel()] = t
    gathered = torch.empty(world * pad, dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(gathered, buf, group=group)
    del buf
    total = int(all_counts.sum())
    full = torch.empty(total, dtype=t.dtype, device=t.device)
    dst, src, ln = gather_tables(assignment, all_counts, pad)
    if t.is_cuda:
        if ctx is None:
            raise ValueError("gather_outputs on CUDA tensors needs the rank's grail Context (segment-copy kernel)")
        torch.cuda.current_stream(t.device).synchronize()          # the all_gather ran on torch's stream
        ctx.copy_segments(full.data_ptr(), gathered.data_ptr(), dst, src, ln, t.element_size())
    else:
        _copy_segments_torch(full, gathered, dst, src, ln)
    return full
