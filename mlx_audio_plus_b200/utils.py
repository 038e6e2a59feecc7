"""Backward-compat re-export seam, mirroring mlx_audio/utils.py:29-38 (callers import the DSP names
from either `dsp` or `utils`; mlx_audio/tests/test_dsp.py:27-34)."""
from .dsp import (  # noqa: F401
    STR_TO_WINDOW_FN,
    bartlett,
    blackman,
    hamming,
    hanning,
    istft,
    mel_filters,
    stft,
)
