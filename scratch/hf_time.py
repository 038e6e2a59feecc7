import sys, torch, numpy as np
sys.path.insert(0, ".")
from mlx_audio_plus_b200.stt.models.qwen3_asr.feature_extractor import WhisperFeatureExtractor
from mlx_audio_plus_b200.stt.models.whisper.audio import log_mel_spectrogram
x = 0.1 * torch.randn(1024, 480000, device="cuda")
fe = WhisperFeatureExtractor(feature_size=128)
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); a = torch.cuda.Event(True); b = torch.cuda.Event(True); a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
print("whisper128 (T,M) generated-mel kernel: %.3f ms" % t(lambda: log_mel_spectrogram(x, n_mels=128)))
print("HF extractor (M,T) run-time tables, incl. pad/mask plumbing: %.3f ms" % t(lambda: fe(x, sampling_rate=16000, padding=True, truncation=False, return_attention_mask=True, return_tensors="cuda")))
from mlx_audio_plus_b200.codec.models.s3tokenizer.utils import log_mel_spectrogram_compat
print("s3tokenizer compat (M,T): %.3f ms" % t(lambda: log_mel_spectrogram_compat(x, 128)))
