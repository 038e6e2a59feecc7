"""Drop-in for mlx_audio/stt/models/whisper/audio.py (constants, pad_or_trim, log_mel_spectrogram).

One fused kernel does what the reference spells as stft -> [:-1] -> abs().square() -> @ filters.T ->
maximum(1e-10).log10() -> maximum(max-8) -> (x+4)/4 (audio.py:73-85).  `padding` zero samples are virtual
(never materialised).  Extension: a (B, L) batch is processed in one launch, each clip with its own max
(as if the reference had been called per clip)."""
from __future__ import annotations

import numpy as np

from ...._arrays import emit
from ...._wrap import as_batch, run_frontend
from .... import _lib as L
from ....dsp import hanning, mel_filters

# hard-coded audio hyperparameters (reference audio.py:14-25)
SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE
N_FRAMES = N_SAMPLES // HOP_LENGTH
N_SAMPLES_PER_TOKEN = HOP_LENGTH * 2
FRAMES_PER_SECOND = SAMPLE_RATE // HOP_LENGTH
TOKENS_PER_SECOND = SAMPLE_RATE // N_SAMPLES_PER_TOKEN


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    """Pad or trim to `length` along `axis` (reference audio.py:27-41); numpy or torch arrays."""
    n = array.shape[axis]
    if n > length:
        sl = [slice(None)] * array.ndim
        sl[axis] = slice(0, length)
        array = array[tuple(sl)]
    if array.shape[axis] < length:
        extra = length - array.shape[axis]
        if type(array).__module__.split(".")[0] == "torch":
            import torch

            shape = list(array.shape)
            shape[axis] = extra
            array = torch.cat([array, torch.zeros(shape, dtype=array.dtype, device=array.device)], dim=axis)
        else:
            pw = [(0, 0)] * array.ndim
            pw[axis] = (0, extra)
            array = np.pad(array, pw)
    return array


def log_mel_spectrogram(audio, n_mels: int = 80, padding: int = 0, *, dtype="float32"):
    """Whisper log-mel features, shape (T, n_mels) float32 (reference audio.py:44-85).
    Extension: dtype="float16" / "bfloat16" writes the encoder's input dtype straight from the fused kernel — the
    `.astype(self.dtype)` of whisper/whisper.py:994-996 without a second pass, bit-identical to casting the float32 result."""
    if isinstance(audio, str):  # audio.py:68-69 `load_audio(audio)`: 16-bit PCM WAVE here, other containers need a decoder
        from ...utils import load_audio

        audio = load_audio(audio)
    ing, was_1d = as_batch(audio)
    fb = mel_filters(SAMPLE_RATE, N_FFT, n_mels, norm="slaney", mel_scale=None)
    out = run_frontend(
        ing, hanning(N_FFT), fb, length=ing.data.shape[1] + max(int(padding), 0),
        n_fft=N_FFT, hop=HOP_LENGTH, center=True, pad_mode="reflect", drop_last=True,
        spec_kind=L.SPEC_POWER, log_kind=L.LOG_LOG10, guard_kind=L.GUARD_MAX, guard_eps=1e-10,
        clamp_kind=L.CLAMP_CLIP_MAX, clamp_value=8.0, affine_add=4.0, affine_div=4.0, out_dtype=str(dtype).split(".")[-1])
    return emit(ing, out[0] if was_1d else out)


def mel_segment(mel, seek: int, segment_size: int, n_frames: int = N_FRAMES, dtype="float16"):
    """The decoder loop's segment builder (reference whisper/whisper.py:990-996),
    ``pad_or_trim(mel[seek : seek + segment_size], N_FRAMES, axis=-2).astype(dtype)``, as one kernel: slice, zero rows up to
    `n_frames`, cast to the encoder dtype ("float16" / "bfloat16" / "float32").  mel: (T, n_mels) or (B, T, n_mels)."""
    from ...._post import rows_pad_cast

    return rows_pad_cast(mel, seek, segment_size, n_frames, dtype)
