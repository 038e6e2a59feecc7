which compute-sanitizer || ls /usr/local/cuda/bin | grep -i sanit
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -q -m gpu -x -k "fast_stft or whisper_length or load_audio_resample or funasr or mel_segment or 16bit or parakeet_and_vocos" 2>&1 | tail -12
echo "exit: $?"
