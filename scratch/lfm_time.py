import sys, torch, numpy as np
sys.path.insert(0, ".")
from mlx_audio_plus_b200.sts.models.lfm_audio.processor import AudioPreprocessor, PreprocessorConfig
from mlx_audio_plus_b200.stt.models.parakeet.audio import PreprocessArgs, log_mel_spectrogram
from mlx_audio_plus_b200.vad.models.sortformer.sortformer import extract_mel_features
from mlx_audio_plus_b200.frontend import _CACHE
x = 0.1 * torch.randn(1024, 480000, device="cuda")
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); a = torch.cuda.Event(True); b = torch.cuda.Event(True); a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
lfm = AudioPreprocessor(PreprocessorConfig(dither=0.0))
print("lfm2 per_feature (valid-frame stats): %.3f ms" % t(lambda: lfm(x)))
lfm0 = AudioPreprocessor(PreprocessorConfig(dither=0.0, normalize="none"))
print("lfm2 no normalisation: %.3f ms" % t(lambda: lfm0(x)))
pa = PreprocessArgs(16000, "per_feature", 0.025, 0.01, "hann", 128, 512, 0.0)
print("parakeet-128 per_feature (all-frame stats, one forward): %.3f ms" % t(lambda: log_mel_spectrogram(x, pa)))
pa80 = PreprocessArgs(16000, "per_feature", 0.025, 0.01, "hann", 80, 512, 0.0)
print("parakeet-80 per_feature: %.3f ms" % t(lambda: log_mel_spectrogram(x, pa80)))
print("sortformer (M,T) 80: %.3f ms" % t(lambda: extract_mel_features(x)))
print(sorted(set(p.kernel_name for p in _CACHE.values())))
