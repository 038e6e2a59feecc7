"""GPU: one cached plan serves concurrent streams and threads (the plan's own statistics scratch was shared
by every caller on a device; the Python layer now hands each call its own workspace — b2a_frontend_call_workspace_bytes)."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle.make_golden import synth  # noqa: E402


def _clips(n, seed):
    x = np.stack([synth(seed + i, 16000 * 20) * (0.05 + 0.3 * i) for i in range(n)])
    x[:, 100000:180000] = 0.0  # digital silence: the per-clip max - 8 clamp rewrites these frames
    return x


def test_shared_plan_on_two_streams_matches_sequential():
    from mlx_audio_plus_b200.stt.models.whisper.audio import log_mel_spectrogram

    xa, xb = torch.from_numpy(_clips(6, 100)).cuda(), torch.from_numpy(_clips(6, 200) * 30.0).cuda()
    ra, rb = log_mel_spectrogram(xa, n_mels=128), log_mel_spectrogram(xb, n_mels=128)  # same cached plan, one after the other
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(20):  # interleaved launches on two streams: statistics of one call must never leak into the other
        with torch.cuda.stream(sa):
            ya = log_mel_spectrogram(xa, n_mels=128)
        with torch.cuda.stream(sb):
            yb = log_mel_spectrogram(xb, n_mels=128)
        torch.cuda.synchronize()
        assert torch.equal(ya, ra) and torch.equal(yb, rb)


def test_shared_plan_from_two_threads_matches_sequential():
    from mlx_audio_plus_b200.stt.models.parakeet.audio import PreprocessArgs, log_mel_spectrogram

    args = PreprocessArgs(sample_rate=16000, normalize="per_feature", window_size=0.025, window_stride=0.01, window="hann",
                          features=80, n_fft=512, dither=0.0)
    xs = [torch.from_numpy(_clips(1, 300 + 7 * i)[0] * (1.0 + 3.0 * i)).cuda() for i in range(4)]
    refs = [log_mel_spectrogram(x, args) for x in xs]
    torch.cuda.synchronize()
    errs = []

    def worker(i):
        try:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for _ in range(25):
                    y = log_mel_spectrogram(xs[i], args)
                    st.synchronize()
                    if not torch.equal(y, refs[i]):
                        errs.append(i)
                        return
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs


def test_host_pcm16_entry_is_bit_identical_to_the_float32_entry():
    """b2a_frontend_forward_host with audio_kind = B2A_PCM_I16: int16 / 32768 on the device (exact) in front of the fused
    kernel — features equal the float32 entry's on the converted samples bit for bit, float32 and float16 rows."""
    from mlx_audio_plus_b200 import _lib as L
    from mlx_audio_plus_b200.dsp import hanning, mel_filters
    from mlx_audio_plus_b200.frontend import FrontendPlan
    from mlx_audio_plus_b200._arrays import ingest

    rng = np.random.default_rng(5)
    pcm = np.clip(np.stack([synth(600 + i, 48017) for i in range(5)]) * 20000.0 + rng.normal(0, 3, (5, 48017)), -32768, 32767).astype(np.int16)
    pcm[2, 10000:30000] = 0
    x = (pcm.astype(np.float32) / np.float32(32768.0)).astype(np.float32)
    for od in ("float32", "float16"):
        plan = FrontendPlan(n_fft=400, hop=160, window=np.asarray(hanning(400)), center=True, pad_mode="reflect", drop_last=True,
                            spec_kind=L.SPEC_POWER, filterbank=np.asarray(mel_filters(16000, 400, 128, norm="slaney", mel_scale=None)),
                            log_kind=L.LOG_LOG10, guard_kind=L.GUARD_MAX, guard_eps=1e-10, clamp_kind=L.CLAMP_CLIP_MAX, clamp_value=8.0,
                            affine_add=4.0, affine_div=4.0, out_dtype=od)
        ref = plan.run(ingest(x, "float32"))
        got = plan.run_host_pcm16(pcm)
        assert got.dtype == ref.dtype and got.shape == ref.shape
        np.testing.assert_array_equal(got, ref)
    with pytest.raises(NotImplementedError):  # the device entries read float32 only
        a = plan._args(0, 10, 10, 10, 1, 0)
        a.audio_kind = L.PCM_I16
        L.check(L.lib.b2a_frontend_forward(plan._h, __import__("ctypes").byref(a), None))
