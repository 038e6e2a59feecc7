"""Drop-in for the STFT pair CosyVoice2's HiFT generator uses (mlx_audio/tts/models/cosyvoice2/hifigan.py:13-19 imports
`stft`, `istft`, `hann_window_periodic` from codec/models/s3gen/hifigan.py; `_stft` / `_istft` at 452-470 forward to them
with the periodic Hann of n_fft taps)."""
from ....codec.models.s3gen.hifigan import hann_window_periodic, istft, stft  # noqa: F401
