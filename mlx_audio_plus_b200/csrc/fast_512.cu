// b200audio — fast fused log-mel kernel, n_fft = 512 instances (Parakeet / Sortformer: 256 threads, 2 CTAs / SM).
#include "fast_fwd.cuh"

namespace b2a {
#ifdef B2A_DEV_400_ONLY  // development builds compile the Whisper variant only
int fast_match_512(const b2a_plan*, const char**) { return 0; }
int fast_launch_512(b2a_plan*, FastState*, FastParams&, cudaStream_t) { return B2A_ERR_UNSUPPORTED; }
#else
namespace {
using Cfg512 = Cfg<16, 16, 160, true, 2>;
B2A_SPECLIST(Cfg512, B2A_SPECS_512)
}  // namespace

int fast_match_512(const b2a_plan* plan, const char** name) { return SpecList<Cfg512>::match(plan, name); }
int fast_launch_512(b2a_plan* plan, FastState* fs, FastParams& p, cudaStream_t st) { return launch<Cfg512>(plan, fs, p, st); }

#endif
}  // namespace b2a
