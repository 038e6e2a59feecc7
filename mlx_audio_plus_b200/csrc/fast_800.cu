// b200audio — fast kernels for two more sizes: n_fft = 800 / hop 200 (the DEFAULTS of dsp.stft, dsp.py:92-104: Nc = 400 =
// 20 x 20, one 10-warp CTA per SM, up to 204 registers) and n_fft = 1024 / hop 320 (Spark BiCodec mel, bicodec.py:20-49).
// No generated mel code for these (run-time filterbank tables); dsp.stft runs on fast_stft_kernel.
#include "fast_fwd.cuh"

namespace b2a {
#ifdef B2A_DEV_400_ONLY  // development builds compile the Whisper variant only
int fast_match_800(const b2a_plan*, const char**) { return 0; }
int fast_launch_800(b2a_plan*, FastState*, FastParams&, cudaStream_t) { return B2A_ERR_UNSUPPORTED; }
int fast_match_1024h320(const b2a_plan*, const char**) { return 0; }
int fast_launch_1024h320(b2a_plan*, FastState*, FastParams&, cudaStream_t) { return B2A_ERR_UNSUPPORTED; }
#else
namespace {
#define B2A_SPECS_NONE(X)
using Cfg800 = Cfg<20, 20, 200, true, 1>;
using Cfg1024h320 = Cfg<32, 16, 320, true, 1>;
B2A_SPECLIST(Cfg800, B2A_SPECS_NONE)
B2A_SPECLIST(Cfg1024h320, B2A_SPECS_NONE)
}  // namespace

int fast_match_800(const b2a_plan* plan, const char** name) { return SpecList<Cfg800>::match(plan, name); }
int fast_launch_800(b2a_plan* plan, FastState* fs, FastParams& p, cudaStream_t st) { return launch<Cfg800>(plan, fs, p, st); }
int fast_match_1024h320(const b2a_plan* plan, const char** name) { return SpecList<Cfg1024h320>::match(plan, name); }
int fast_launch_1024h320(b2a_plan* plan, FastState* fs, FastParams& p, cudaStream_t st) { return launch<Cfg1024h320>(plan, fs, p, st); }

#endif
}  // namespace b2a
