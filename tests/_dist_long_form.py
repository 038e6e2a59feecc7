"""Worker of tests/test_gpu_multi.py: ONE rank of a real multi-GPU long-form run (one process per GPU, NCCL).
Each rank holds only its sample slice (halo included) of the signal, runs parallel.long_form_features — partial on its
frame range, the NCCL all-reduce of the statistics, finalize — gathers the shards with parallel.gather_features and
rank 0 writes the result.  Launched by the test with RANK / WORLD_SIZE / MASTER_* in the environment."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from mlx_audio_plus_b200 import _lib as L
    from mlx_audio_plus_b200.dsp import hanning, mel_filters
    from mlx_audio_plus_b200.frontend import FrontendPlan
    from mlx_audio_plus_b200.parallel import clip_shard, frame_shards, gather_features, long_form_features, num_frames
    from oracle.make_golden import synth

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    out_path, length = sys.argv[1], int(sys.argv[2])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    try:
        x = synth(41, length)
        x[length // 3 : length // 3 + 30000] = 0  # digital silence: the global max - 8 clamp acts across shard boundaries
        res = {}
        # ---- Parakeet: per-feature mean / std over ALL frames -> all_reduce(SUM) of 2 * 80 float64 -------------
        plan = FrontendPlan(n_fft=512, hop=160, window=np.asarray(hanning(400)), preemph=0.97, spec_kind=L.SPEC_POWER,
                            filterbank=np.asarray(mel_filters(16000, 512, 80, norm="per_feature", mel_scale=None)),
                            log_kind=L.LOG_LN, guard_kind=L.GUARD_ADD, guard_eps=1e-5, norm_kind=L.NORM_PER_FEATURE,
                            norm_ddof=0, norm_eps=1e-5)
        sh = frame_shards(length, 512, 160, world, preemph=True)[rank]
        xs = torch.from_numpy(x[sh.sample_lo : sh.sample_hi].copy()).to(dev)  # this rank never sees the rest of the signal
        y = long_form_features(plan, xs, sh, length=length, global_frames=num_frames(length, 512, 160))
        res["parakeet"] = gather_features(y)
        # ---- Whisper long form: ONE max over the whole file -> all_reduce(MAX) of 1 float ----------------------
        planw = FrontendPlan(n_fft=400, hop=160, window=np.asarray(hanning(400)), drop_last=True, spec_kind=L.SPEC_POWER,
                             filterbank=np.asarray(mel_filters(16000, 400, 128, norm="slaney", mel_scale=None)),
                             log_kind=L.LOG_LOG10, guard_kind=L.GUARD_MAX, guard_eps=1e-10, clamp_kind=L.CLAMP_CLIP_MAX,
                             clamp_value=8.0, affine_add=4.0, affine_div=4.0)
        shw = frame_shards(length, 400, 160, world, drop_last=True)[rank]
        xw = torch.from_numpy(x[shw.sample_lo : shw.sample_hi].copy()).to(dev)
        yw = long_form_features(planw, xw, shw, length=length, global_frames=num_frames(length, 400, 160, True, True))
        res["whisper"] = gather_features(yw)
        # ---- clip sharding: each rank featurises its clips, no collective on the data path ---------------------
        clips = np.stack([synth(300 + i, 48000) * (0.2 + 0.1 * i) for i in range(7)])
        c0, c1 = clip_shard(7, world, rank)  # 7 clips: ragged over 2 / 4 ranks, and rank 7 of 8 owns NO clip
        from mlx_audio_plus_b200.stt.models.whisper.audio import log_mel_spectrogram
        if c1 > c0:
            yc = log_mel_spectrogram(torch.from_numpy(clips[c0:c1]).to(dev), n_mels=80)
        else:
            yc = torch.empty((0, 300, 80), dtype=torch.float32, device=dev)
        res["clips"] = gather_features(yc)
        torch.cuda.synchronize()
        if rank == 0:
            np.savez(out_path, **{k: v.cpu().numpy() for k, v in res.items()}, world=np.int64(dist.get_world_size()),
                     backend=np.bytes_(dist.get_backend()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
