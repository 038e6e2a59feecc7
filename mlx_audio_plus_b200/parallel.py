"""Data-parallel sharding of the hot path across the GPUs of one box (one process per GPU).

The path shards with NO data-path collective:
  * batches of clips  -> contiguous clip ranges per rank (`clip_shard`);
  * one long signal   -> contiguous FRAME ranges per rank (`frame_shards`); each rank reads its sample slice
    plus a halo of n_fft - hop samples (+1 with pre-emphasis); centre padding applies at the global ends only.
The only exchange step is the tiny cross-frame statistic of the long-form case — Whisper's global max
(1 float, MAX) or Parakeet's per-feature sum / sum of squares (2*M float64, SUM) — done with
torch.distributed (NCCL over NVLink on GPUs, gloo on CPU), followed by a local clamp / normalise.
`gather_features` is the optional all-gather of feature shards and is never part of throughput numbers.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np


def clip_shard(n_clips: int, world: int, rank: int):
    """Contiguous, balanced clip range [start, stop) for `rank`."""
    base, rem = divmod(n_clips, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


@dataclass(frozen=True)
class FrameShard:
    rank: int
    frame_begin: int   # first frame owned (global index)
    frame_count: int
    sample_lo: int     # slice [sample_lo, sample_hi) of the UNPADDED signal the rank must hold
    sample_hi: int


def num_frames(length: int, n_fft: int, hop: int, center: bool = True, drop_last: bool = False) -> int:
    """Frames the reference produces (dsp.py:118-136; reflect/constant centre pad of n_fft//2)."""
    p = n_fft // 2 if center else 0
    pl = min(p, max(length - 1, 0)) if center else 0  # reflect slices truncate when length <= p
    padded = length + 2 * pl
    T = 1 + (padded - n_fft) // hop if padded >= n_fft else 0
    return T - (1 if drop_last else 0)


def frame_shards(length: int, n_fft: int, hop: int, world: int, *, center: bool = True, drop_last: bool = False,
                 preemph: bool = False) -> List[FrameShard]:
    """Split the frames of one signal into `world` contiguous ranges and compute each rank's sample slice."""
    T = num_frames(length, n_fft, hop, center, drop_last)
    p = n_fft // 2 if center else 0
    out = []
    for r in range(world):
        t0, t1 = clip_shard(T, world, r)
        if t1 <= t0:
            out.append(FrameShard(r, t0, 0, 0, 0))
            continue
        # padded coordinates touched: [t0*hop, (t1-1)*hop + n_fft); source = padded - p, reflected at the ends
        lo = t0 * hop - p
        hi = (t1 - 1) * hop + n_fft - p
        # reflection folds indices back inside [0, length): |lo| <= p and 2*length-2-hi' stay within the end frames
        s_lo = max(0, min(lo, length - 1)) if lo >= 0 else 0
        s_hi = min(length, hi) if hi <= length else length
        if lo < 0:  # left reflect reads x[1 .. -lo]
            s_hi = max(s_hi, min(length, -lo + 1))
        if hi > length:  # right reflect reads x[2*length-2-(hi-1) ..]
            s_lo = min(s_lo, max(0, 2 * length - 1 - hi))
        if preemph and s_lo > 0:
            s_lo -= 1  # y[n] needs x[n-1]
        out.append(FrameShard(r, t0, t1 - t0, s_lo, s_hi))
    return out


def reduce_stats(clip_max, feat_sums, group=None, *, need_max=True, need_sums=True):
    """All-reduce the cross-frame statistics of a frame-sharded signal (MAX for the clip maximum, SUM for the
    per-feature sums).  Works on torch CUDA tensors (NCCL) and CPU tensors (gloo).  `need_max` / `need_sums` skip the
    exchange a plan does not use (a clamping front-end needs only the maximum, a normalising one only the sums): one
    collective per step instead of two — a frame-sharded step is latency-bound."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return clip_max, feat_sums
    if need_max:
        dist.all_reduce(clip_max, op=dist.ReduceOp.MAX, group=group)
    if need_sums:
        dist.all_reduce(feat_sums, op=dist.ReduceOp.SUM, group=group)
    return clip_max, feat_sums


def long_form_features(plan, signal_slice, shard: FrameShard, *, length: int, global_frames: int, group=None,
                       valid_length=None, pad_value=0.0, out=None):
    """One rank's part of a frame-sharded long-form featurisation.  `signal_slice` is the torch CUDA tensor
    holding samples [shard.sample_lo, shard.sample_hi) (1-D).  Returns this rank's (frames, n_out) features,
    clamped / normalised with the GLOBAL statistics (whole-file statistics as parakeet/audio.py:66-69 and
    whisper/audio.py:83 compute them).  `out`: optional preallocated (1, frames, n_out) float32 buffer."""
    import torch

    from . import _lib as L

    x = signal_slice.reshape(1, -1).contiguous()
    if out is None:
        out = torch.empty(plan.out_shape(1, shard.frame_count), dtype=torch.float32, device=x.device)
    clip_max, feat_sums = plan.stats_tensors(1, x.device)
    if shard.frame_count > 0:
        plan.partial(x, out, clip_max, feat_sums, length=length, sample_offset=shard.sample_lo,
                     frame_begin=shard.frame_begin, frame_count=shard.frame_count, valid_length=valid_length,
                     pad_value=pad_value)
    reduce_stats(clip_max, feat_sums, group, need_max=plan.desc.clamp_kind != L.CLAMP_NONE,
                 need_sums=plan.desc.norm_kind != L.NORM_NONE)
    if shard.frame_count > 0:
        plan.finalize(out, clip_max, feat_sums, global_frames=global_frames)
    return out[0]


def gather_features(local, group=None):
    """Optional all-gather of feature shards along the frame / clip axis (axis 0) to every rank (reported separately from
    throughput: the path itself needs no collective).  Frame ranges differ by at most one frame between ranks, so the
    shard lengths are exchanged first and shorter shards travel zero-padded to the longest."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    counts = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    top = max(counts)
    send = local.contiguous()
    if send.shape[0] < top:
        pad = torch.zeros((top - send.shape[0],) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        send = torch.cat([send, pad], dim=0)
    parts = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(parts, send, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
