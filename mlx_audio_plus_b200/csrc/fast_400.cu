// b200audio — fast fused log-mel kernel, n_fft = 400 instances (Whisper / Voxtral-RT / S3Tokenizer / FunASR: 320 threads, 2 CTAs / SM).
#include "fast_fwd.cuh"

namespace b2a {
namespace {
using Cfg400 = Cfg<20, 10, 160, false, 2>;
B2A_SPECLIST(Cfg400, B2A_SPECS_400)
}  // namespace

int fast_match_400(const b2a_plan* plan, const char** name) { return SpecList<Cfg400>::match(plan, name); }
int fast_launch_400(b2a_plan* plan, FastState* fs, FastParams& p, cudaStream_t st) { return launch<Cfg400>(plan, fs, p, st); }

}  // namespace b2a
