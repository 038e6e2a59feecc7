"""mlx_audio_plus_b200 — B200-native (sm_100a CUDA) drop-in for the STFT / log-mel / iSTFT hot path of
mlx-audio-plus (`mlx_audio/dsp.py` + each model's `log_mel_spectrogram`).  See DESIGN.md / INTEGRATION.md."""
from .version import __version__  # noqa: F401
