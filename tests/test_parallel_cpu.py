"""CPU (gloo, world_size 2): the host-side sharding logic of mlx_audio_plus_b200/parallel.py.

Each rank featurises ITS frame range of one long signal from ITS sample slice (halo included) with the
oracle standing in for the kernel, the ranks all-reduce the statistics through parallel.reduce_stats over gloo,
finalize locally, and the concatenation must equal the oracle run on the whole signal.  Clip sharding is
checked the same way."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mlx_audio_plus_b200.parallel import clip_shard, frame_shards, gather_features, num_frames, reduce_stats
from oracle import dsp_oracle as D
from oracle import wrappers_oracle as W
from oracle.make_golden import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _raw_logmel_frames(x_slice, sample_lo, length, shard, n_fft, hop, fb, preemph, log):
    """Oracle stand-in for `plan.partial`: un-normalised log-mel of the shard's frames, computed from the
    slice only (global index map -> local)."""
    idx = D.frame_indices(length, n_fft, hop, True, "reflect")[shard.frame_begin : shard.frame_begin + shard.frame_count]
    x = np.asarray(x_slice, np.float32)
    if preemph:
        prev = np.where(idx - 1 >= 0, idx - 1, 0)
        fr = x[idx - sample_lo] - np.where(idx > 0, np.float32(preemph) * x[prev - sample_lo], np.float32(0))
        fr = fr.astype(np.float32)
    else:
        fr = x[idx - sample_lo]
    w = D._fit_window(D.hanning(400), n_fft)
    spec = np.fft.rfft(fr * w).astype(np.complex64)
    mel = (np.abs(spec) ** 2).astype(np.float32) @ fb.T
    return log(mel)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- Parakeet-style long form: per-feature normalisation needs the SUM all-reduce -------------------
        L, n_fft, hop, M = 40000, 512, 160, 80
        x = synth(77, L)
        fb = D.mel_filters(16000, n_fft, M, norm="per_feature", mel_scale=None)
        shards = frame_shards(L, n_fft, hop, world, preemph=True)
        sh = shards[rank]
        xs = x[sh.sample_lo : sh.sample_hi]  # this rank only ever touches its slice
        raw = _raw_logmel_frames(xs, sh.sample_lo, L, sh, n_fft, hop, fb, 0.97, lambda m: np.log(m + np.float32(1e-5)))
        clip_max = torch.tensor([raw.max()], dtype=torch.float32)
        sums = torch.tensor(np.stack([raw.astype(np.float64).sum(0), (raw.astype(np.float64) ** 2).sum(0)], -1))[None]
        reduce_stats(clip_max, sums)
        T = num_frames(L, n_fft, hop)
        mean = sums[0, :, 0].numpy() / T
        std = np.sqrt(np.maximum(sums[0, :, 1].numpy() / T - mean**2, 0))
        part = (raw - mean.astype(np.float32)) / (std.astype(np.float32) + np.float32(1e-5))
        # ---- Whisper-style long form: global max needs the MAX all-reduce -----------------------------------
        fbw = D.mel_filters(16000, 400, 80, norm="slaney", mel_scale=None)
        shw = frame_shards(L, 400, hop, world, drop_last=True)[rank]
        xw = x[shw.sample_lo : shw.sample_hi]
        raww = _raw_logmel_frames(xw, shw.sample_lo, L, shw, 400, hop, fbw, 0.0, lambda m: np.log10(np.maximum(m, np.float32(1e-10))))
        mx = torch.tensor([raww.max()], dtype=torch.float32)
        reduce_stats(mx, torch.zeros(1, 80, 2, dtype=torch.float64))
        partw = (np.maximum(raww, mx.item() - np.float32(8.0)) + np.float32(4.0)) / np.float32(4.0)
        # ---- clip sharding ----------------------------------------------------------------------------------
        c0, c1 = clip_shard(7, world, rank)
        # ---- ragged all-gather of the shards (frame counts differ by one between the ranks) -------------------
        full = gather_features(torch.from_numpy(np.ascontiguousarray(part))).numpy()
        q.put((rank, sh.frame_begin, part, shw.frame_begin, partw, (c0, c1), full))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_frame_range_sharding_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=150) for _ in range(world)])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    L = 40000
    x = synth(77, L)
    pa = W.PreprocessArgs(16000, "per_feature", 0.025, 0.01, "hann", 80, 512, 1e-5)
    ref = W.parakeet_log_mel(x, pa)[0]
    got = np.concatenate([r[2] for r in res])
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 5e-4
    refw = W.whisper_log_mel(x, 80)
    gotw = np.concatenate([r[4] for r in res])
    assert gotw.shape == refw.shape and np.abs(gotw - refw).max() <= 1e-5
    assert [r[5] for r in res] == [(0, 4), (4, 7)]
    assert res[0][2].shape[0] != res[1][2].shape[0]  # 251 frames over 2 ranks: ragged
    for r in res:  # every rank ends up with the whole feature matrix
        np.testing.assert_array_equal(r[6], got)


@pytest.mark.parametrize("L,n_fft,hop,center,pre", [(57600, 512, 160, True, True), (48000, 400, 160, True, False),
                                                    (4000, 400, 160, True, False), (9999, 1024, 256, False, False),
                                                    (480000, 400, 160, True, False)])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_frame_shards_cover_exactly_what_each_rank_reads(L, n_fft, hop, center, pre, world):
    shards = frame_shards(L, n_fft, hop, world, center=center, preemph=pre)
    idx = D.frame_indices(L, n_fft, hop, center, "reflect")
    assert sum(s.frame_count for s in shards) == idx.shape[0] == num_frames(L, n_fft, hop, center)
    nxt = 0
    for s in shards:
        assert s.frame_begin == nxt
        nxt += s.frame_count
        if s.frame_count == 0:
            continue
        need = idx[s.frame_begin : s.frame_begin + s.frame_count]
        lo, hi = int(need.min()), int(need.max())
        assert s.sample_lo <= max(0, lo - (1 if pre else 0)) and s.sample_hi > hi
        # halo is bounded: a rank never needs more than its frames' span plus n_fft of context
        assert s.sample_hi - s.sample_lo <= s.frame_count * hop + 2 * n_fft + 1


def test_clip_shard_partition():
    for n in (1, 7, 4096):
        for w in (1, 2, 3, 8):
            parts = [clip_shard(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
