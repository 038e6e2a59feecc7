// b200audio — fast fused log-mel kernel, n_fft = 1024 instances (Vocos / Qwen3-TTS mel: 512 threads, 1 CTA / SM).
#include "fast_fwd.cuh"

namespace b2a {
#ifdef B2A_DEV_400_ONLY  // development builds compile the Whisper variant only
int fast_match_1024(const b2a_plan*, const char**) { return 0; }
int fast_launch_1024(b2a_plan*, FastState*, FastParams&, cudaStream_t) { return B2A_ERR_UNSUPPORTED; }
#else
namespace {
using Cfg1024 = Cfg<32, 16, 256, true, 1>;
B2A_SPECLIST(Cfg1024, B2A_SPECS_1024)
}  // namespace

int fast_match_1024(const b2a_plan* plan, const char** name) { return SpecList<Cfg1024>::match(plan, name); }
int fast_launch_1024(b2a_plan* plan, FastState* fs, FastParams& p, cudaStream_t st) { return launch<Cfg1024>(plan, fs, p, st); }

#endif
}  // namespace b2a
