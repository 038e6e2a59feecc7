"""GPU: frame-range sharding through the split C-ABI (partial -> reduce -> finalize).  The ranks are emulated
on ONE GPU as sequential launches over each rank's own slice; statistics are merged exactly as the NCCL
all-reduce would (MAX / SUM)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from mlx_audio_plus_b200 import _lib as L  # noqa: E402
from mlx_audio_plus_b200.dsp import hanning, mel_filters  # noqa: E402
from mlx_audio_plus_b200.frontend import FrontendPlan  # noqa: E402
from mlx_audio_plus_b200.parallel import frame_shards, num_frames  # noqa: E402
from oracle import wrappers_oracle as W  # noqa: E402
from oracle.make_golden import synth  # noqa: E402


def _run_sharded(make_plan, x, length, world, *, preemph, drop_last):
    plan = make_plan()
    shards = frame_shards(length, plan.n_fft, plan.hop, world, preemph=preemph, drop_last=drop_last)
    outs, maxes, sums = [], [], []
    for sh in shards:
        plan = make_plan()  # one plan per (emulated) rank, as in a real one-process-per-GPU run
        xs = torch.from_numpy(x[sh.sample_lo : sh.sample_hi].copy()).cuda()[None]
        out = torch.empty(plan.out_shape(1, sh.frame_count), dtype=torch.float32, device="cuda")
        cm, fs = plan.stats_tensors(1, out.device)
        plan.partial(xs, out, cm, fs, length=length, sample_offset=sh.sample_lo, frame_begin=sh.frame_begin,
                     frame_count=sh.frame_count)
        outs.append((out, plan))
        maxes.append(cm)
        sums.append(fs)
    gmax = torch.stack(maxes).max(0).values  # == all_reduce(MAX)
    gsum = torch.stack(sums).sum(0)  # == all_reduce(SUM)
    T = num_frames(length, plan.n_fft, plan.hop, True, drop_last)
    res = []
    for out, plan in outs:
        plan.finalize(out, gmax.clone(), gsum.clone(), global_frames=T)
        res.append(out[0].cpu().numpy())
    return np.concatenate(res)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_whisper_long_form_sharded(world):
    x = synth(5, 160000)
    x[60000:90000] = 0  # silence: the global max-8 clamp is active across shard boundaries
    def plan():
        return FrontendPlan(n_fft=400, hop=160, window=np.asarray(hanning(400)), drop_last=True, spec_kind=L.SPEC_POWER,
                            filterbank=np.asarray(mel_filters(16000, 400, 80, norm="slaney", mel_scale=None)),
                            log_kind=L.LOG_LOG10, guard_kind=L.GUARD_MAX, guard_eps=1e-10, clamp_kind=L.CLAMP_CLIP_MAX,
                            clamp_value=8.0, affine_add=4.0, affine_div=4.0)

    got = _run_sharded(plan, x, len(x), world, preemph=False, drop_last=True)
    ref = W.whisper_log_mel(x, 80)
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-4


@pytest.mark.parametrize("world", [2, 8])
def test_parakeet_long_form_sharded(world):
    x = synth(6, 200000)
    def plan():
        return FrontendPlan(n_fft=512, hop=160, window=np.asarray(hanning(400)), preemph=0.97, spec_kind=L.SPEC_POWER,
                            filterbank=np.asarray(mel_filters(16000, 512, 80, norm="per_feature", mel_scale=None)),
                            log_kind=L.LOG_LN, guard_kind=L.GUARD_ADD, guard_eps=1e-5, norm_kind=L.NORM_PER_FEATURE,
                            norm_ddof=0, norm_eps=1e-5)

    got = _run_sharded(plan, x, len(x), world, preemph=True, drop_last=False)
    pa = W.PreprocessArgs(16000, "per_feature", 0.025, 0.01, "hann", 80, 512, 1e-5)
    ref = W.parakeet_log_mel(x, pa)[0]
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 5e-4
