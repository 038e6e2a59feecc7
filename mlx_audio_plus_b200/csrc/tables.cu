// b200audio — host-side tables and geometry: windows, mel filterbank, framing arithmetic.
// Mirrors (does not copy) mlx_audio/dsp.py:33-88 (windows), 223-296 (mel_filters), 118-136 (framing).
#include <math.h>
#include <stdarg.h>

#include "common.cuh"

namespace b2a {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return B2A_ERR_CUDA;
}

}  // namespace b2a

extern "C" {

int b2a_version(void) { return B2A_VERSION; }

const char* b2a_last_error(void) { return b2a::g_err; }

int b2a_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// dsp.py:33-79: each tap is evaluated in float64 (Python math.cos) and rounded to float32 once.
int b2a_window(int kind, int size, int periodic, float* out) {
  if (size <= 0 || !out) {
    b2a::set_error("b2a_window: invalid size %d", size);
    return B2A_ERR_INVALID_ARG;
  }
  const double D = periodic ? (double)size : (double)(size - 1);
  const double two_pi = 2 * M_PI, four_pi = 4 * M_PI;
  for (int n = 0; n < size; ++n) {
    double v;
    switch (kind) {
      case B2A_WIN_HANN:
        v = 0.5 * (1 - cos(two_pi * n / D));
        break;
      case B2A_WIN_HAMMING:
        v = 0.54 - 0.46 * cos(two_pi * n / D);
        break;
      case B2A_WIN_BLACKMAN:
        v = 0.42 - 0.5 * cos(two_pi * n / D) + 0.08 * cos(four_pi * n / D);
        break;
      case B2A_WIN_BARTLETT:
        v = 1 - 2 * fabs(n - D / 2) / D;
        break;
      default:
        b2a::set_error("Unknown window function: kind=%d", kind);
        return B2A_ERR_UNKNOWN_WINDOW;
    }
    out[n] = (float)v;
  }
  return B2A_OK;
}

// float32 linspace in MLX's form (1-t)*start + t*stop, t = i/(num-1)
static void linspace_f32(double start, double stop, int num, std::vector<float>& v) {
  v.resize(num);
  if (num == 1) {
    v[0] = (float)start;
    return;
  }
  const float a = (float)start, b = (float)stop, d = (float)(num - 1);
  for (int i = 0; i < num; ++i) {
    volatile float t = (float)i / d;  // volatile: keep every intermediate rounded to fp32
    volatile float u = 1.0f - t;
    volatile float p = u * a;
    volatile float q = t * b;
    v[i] = p + q;
  }
}

int b2a_mel_filters(int sample_rate, int n_fft, int n_mels, double f_min, double f_max, int norm_slaney,
                    int scale_htk, float* out) {
  if (sample_rate <= 0 || n_fft <= 0 || n_mels <= 0 || !out) {
    b2a::set_error("b2a_mel_filters: invalid argument");
    return B2A_ERR_INVALID_ARG;
  }
  const int F = n_fft / 2 + 1;
  if (!(f_max > 0)) f_max = sample_rate / 2.0;  // dsp.py:264 `f_max or sample_rate / 2`
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
  const double logstep = log(6.4) / 27.0;
  auto hz_to_mel = [&](double f) -> double {  // dsp.py:233-245, float64 scalars
    if (scale_htk) return 2595.0 * log10(1.0 + f / 700.0);
    if (f >= min_log_hz) return min_log_mel + log(f / min_log_hz) / logstep;
    return f / f_sp;
  };
  std::vector<float> bins, grid;
  linspace_f32(0.0, (double)(sample_rate / 2), F, bins);  // integer sr//2, dsp.py:269
  linspace_f32(hz_to_mel(f_min), hz_to_mel(f_max), n_mels + 2, grid);
  std::vector<float> f_pts(n_mels + 2);
  for (int j = 0; j < n_mels + 2; ++j) {  // dsp.py:247-262, float32 arrays
    const float m = grid[j];
    if (scale_htk) {
      volatile float e = m / 2595.0f;
      volatile float pw = (float)pow(10.0, (double)e);  // correctly rounded float32 10**x
      volatile float d = pw - 1.0f;
      f_pts[j] = 700.0f * d;
    } else {
      if (m >= (float)min_log_mel) {
        volatile float d = m - (float)min_log_mel;
        volatile float a = (float)logstep * d;
        volatile float ex = (float)exp((double)a);  // correctly rounded float32 exp
        f_pts[j] = (float)min_log_hz * ex;
      } else {
        f_pts[j] = (float)f_sp * m;
      }
    }
  }
  for (int j = 0; j < n_mels; ++j) {
    volatile float w_lo = f_pts[j + 1] - f_pts[j];
    volatile float w_hi = f_pts[j + 2] - f_pts[j + 1];
    volatile float span = f_pts[j + 2] - f_pts[j];
    const float enorm = norm_slaney ? 2.0f / span : 1.0f;
    for (int f = 0; f < F; ++f) {
      volatile float d0 = f_pts[j] - bins[f];      // slopes[:, j]
      volatile float d2 = f_pts[j + 2] - bins[f];  // slopes[:, j+2]
      volatile float down = (-d0) / w_lo;
      volatile float up = d2 / w_hi;
      float v = fminf(down, up);
      v = fmaxf(0.0f, v);
      if (norm_slaney) {
        volatile float s = v * enorm;
        v = s;
      }
      out[(size_t)j * F + f] = v;
    }
  }
  return B2A_OK;
}

int b2a_stft_geometry(int64_t length, int n_fft, int hop, int center, int pad_mode, int64_t* padded_len,
                      int64_t* num_frames) {
  if (n_fft <= 0 || hop <= 0 || length < 0) {
    b2a::set_error("stft: invalid n_fft=%d hop=%d length=%lld", n_fft, hop, (long long)length);
    return B2A_ERR_INVALID_ARG;
  }
  if (center && pad_mode != B2A_PAD_REFLECT && pad_mode != B2A_PAD_CONSTANT) {
    b2a::set_error("Invalid pad_mode %d", pad_mode);
    return B2A_ERR_PAD_MODE;
  }
  b2a::Geometry g = b2a::make_geometry(length, n_fft, hop, center, pad_mode);
  // python floor division: 1 + (padded - n_fft)//hop <= 0  <=>  padded < n_fft
  if (padded_len) *padded_len = g.padded_len;
  if (num_frames) *num_frames = g.num_frames;
  if (g.num_frames <= 0) {
    b2a::set_error("Input is too short (length=%lld) for n_fft=%d with hop_length=%d and center=%s.",
                   (long long)g.padded_len, n_fft, hop, center ? "True" : "False");
    return B2A_ERR_TOO_SHORT;
  }
  return B2A_OK;
}

int64_t b2a_frame_source_index(int64_t length, int n_fft, int hop, int center, int pad_mode, int64_t t, int k) {
  b2a::Geometry g = b2a::make_geometry(length, n_fft, hop, center, pad_mode);
  return b2a::source_index(g, pad_mode, t * hop + k);
}

// dsp.py:184,211-215 / 390,410-415
int b2a_istft_geometry(int64_t num_frames, int n_fft, int hop, int center, int64_t length, int64_t* ola_len,
                       int64_t* out_start, int64_t* out_len) {
  if (num_frames <= 0 || n_fft <= 0 || hop <= 0) {
    b2a::set_error("istft: invalid num_frames=%lld n_fft=%d hop=%d", (long long)num_frames, n_fft, hop);
    return B2A_ERR_INVALID_ARG;
  }
  const int64_t t = (num_frames - 1) * hop + n_fft;
  int64_t start = 0, len = t;
  if (center && length < 0) {
    // rec[win//2 : -win//2]  — python: -win//2 == -(ceil(win/2))
    start = n_fft / 2;
    int64_t stop = t - (n_fft + 1) / 2;
    if (start > t) start = t;
    len = stop > start ? stop - start : 0;
  }
  if (length >= 0) {  // rec[:length] — the centre pad is NOT stripped
    start = 0;
    len = length < t ? length : t;
  }
  if (ola_len) *ola_len = t;
  if (out_start) *out_start = start;
  if (out_len) *out_len = len;
  return B2A_OK;
}

}  // extern "C"
