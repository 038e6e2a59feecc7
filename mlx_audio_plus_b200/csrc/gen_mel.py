"""Build-time generator of csrc/mel_gen.cuh: straight-line mel-projection code for the NAMED filterbanks.

A mel filterbank is a banded matrix (a frequency bin feeds at most two adjacent triangles), but its band
structure is run-time data, and with run-time tables the projection costs ~6 instructions per tap (address
arithmetic, length dispatch, weight loads).  For the filterbanks the reference's wrappers actually use
(SURVEY.md App. A "wrapper parameter matrix") this script bakes the structure into code: per warp a straight
line of `p = P[k]; acc_m = fmaf(p, W, acc_m)` with the bin offset and the weight as immediates — one shared-
memory load per BIN (shared by the two rows it feeds) and one FFMA per tap.

The filterbank itself is NOT recomputed here: it comes from the product's own host routine `b2a_mel_filters`
(csrc/tables.cu, compiled on the fly into a scratch shared object), so the generated tables are bit-identical
to what a plan receives at run time.  At plan creation fast_fwd.cu compares the caller's filterbank with the
generated one bit for bit; any other filterbank takes the run-time-table path of the same kernel.

Run: python -m mlx_audio_plus_b200.csrc.gen_mel   (also invoked by csrc/build.py when tables.cu changes)
"""
import ctypes
import os
import shutil
import struct
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "mel_gen.cuh")

# name, sample_rate, n_fft, n_mels, f_min, f_max, norm_slaney, scale_htk, warps of the fast kernel variant
SPECS = [
    # Whisper / GLM-ASR / Smart-Turn / S3Tokenizer-compat (whisper/audio.py:76; mel_scale=None -> Slaney)
    ("whisper80", 16000, 400, 80, 0.0, 0.0, 1, 0, 10),
    # Whisper large-v3 / Voxtral-RT / S3Tokenizer (voxtral_realtime/audio.py:19-33: f_max=8000 == sr/2)
    ("whisper128", 16000, 400, 128, 0.0, 0.0, 1, 0, 10),
    # FunASR (funasr/audio.py:32-81: htk scale, slaney norm)
    ("funasr80", 16000, 400, 80, 0.0, 0.0, 1, 1, 10),
    # Parakeet / NeMo (parakeet/audio.py:59-61: norm="per_feature" -> no Slaney normalisation)
    ("parakeet80", 16000, 512, 80, 0.0, 0.0, 0, 0, 8),
    ("parakeet128", 16000, 512, 128, 0.0, 0.0, 0, 0, 8),
    # Sortformer / LFM2 style (slaney scale + slaney norm on n_fft=512)
    ("nemo_slaney80", 16000, 512, 80, 0.0, 0.0, 1, 0, 8),
    ("nemo_slaney128", 16000, 512, 128, 0.0, 0.0, 1, 0, 8),
    # Vocos / IndexTTS (vocos/mel.py:26: htk, no norm, 24 kHz)
    ("vocos100", 24000, 1024, 100, 0.0, 0.0, 0, 1, 16),
    # Qwen3-TTS speaker mel (qwen3_tts.py:33-90: slaney/slaney, f_max 12000 == sr/2)
    ("qwen3tts128", 24000, 1024, 128, 0.0, 12000.0, 1, 0, 16),
    # Hugging Face WhisperFeatureExtractor (Qwen3-ASR / Qwen3-ForcedAligner, qwen3_asr.py:800-846): transformers'
    # mel_filter_bank(norm="slaney", mel_scale="slaney") evaluated in float64 and rounded once — NOT bit-identical to
    # dsp.mel_filters' float32 arithmetic, hence its own specs (filterbank from _hf_filterbank below)
    ("hf_whisper80", 16000, 400, 80, 0.0, 8000.0, 1, 0, 10),
    ("hf_whisper128", 16000, 400, 128, 0.0, 8000.0, 1, 0, 10),
]


def _nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def _tables_lib(tmp):
    so = os.path.join(tmp, "libb2a_tables.so")
    cmd = [_nvcc(), "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-o", so,
           os.path.join(HERE, "tables.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed on tables.cu:\n" + r.stdout + r.stderr)
    lib = ctypes.CDLL(so)
    lib.b2a_mel_filters.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                    ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_float)]
    lib.b2a_mel_filters.restype = ctypes.c_int
    return lib


def _bits(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


def _hexf(x):
    # exact C++17 hexadecimal float literal of a float32 value
    return float(x).hex() + "f"


def _filterbank(lib, sr, n_fft, M, fmin, fmax, norm, htk):
    F = n_fft // 2 + 1
    buf = (ctypes.c_float * (M * F))()
    rc = lib.b2a_mel_filters(sr, n_fft, M, fmin, fmax, norm, htk, buf)
    if rc != 0:
        raise RuntimeError("b2a_mel_filters failed")
    return [[buf[m * F + f] for f in range(F)] for m in range(M)], F


def _hf_filterbank(sr, n_fft, M, fmin, fmax):
    """transformers.audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney") in float64, rounded to float32 — the same
    restatement as stt/models/qwen3_asr/feature_extractor.py::mel_filter_bank_slaney (pinned against transformers itself in
    tests/test_oracle_golden.py); rows = mel filters."""
    import numpy as np

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f >= 1000.0, 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * (27.0 / np.log(6.4)), 3.0 * f / 200.0)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        return np.where(m >= 15.0, 1000.0 * np.exp((np.log(6.4) / 27.0) * (m - 15.0)), 200.0 * m / 3.0)

    F = n_fft // 2 + 1
    ff = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), M + 2))
    fft = np.linspace(0, sr // 2, F)
    d = np.diff(ff)
    sl = np.expand_dims(ff, 0) - np.expand_dims(fft, 1)
    fb = np.maximum(np.zeros(1), np.minimum(-sl[:, :-2] / d[:-1], sl[:, 2:] / d[1:]))
    fb = (fb * np.expand_dims(2.0 / (ff[2 : M + 2] - ff[:M]), 0)).astype(np.float32).T  # (M, F)
    return [[float(v) for v in row] for row in fb], F


def _emit_spec(lib, spec):
    name, sr, n_fft, M, fmin, fmax, norm, htk, NW = spec
    if name.startswith("hf_"):
        fb, F = _hf_filterbank(sr, n_fft, M, fmin, fmax)
    else:
        fb, F = _filterbank(lib, sr, n_fft, M, fmin, fmax, norm, htk)
    start, length = [], []
    for m in range(M):
        nz = [f for f in range(F) if fb[m][f] != 0.0]
        if nz:
            start.append(nz[0])
            length.append(nz[-1] - nz[0] + 1)
        else:
            start.append(0)
            length.append(0)
    wbits = []
    for m in range(M):
        wbits += [_bits(fb[m][start[m] + j]) for j in range(length[m])]
    nnz = len(wbits)
    # contiguous blocks of row QUADS per warp (a quad is stored with one STS.128), balanced by tap + load cost
    assert M % 4 == 0, "the generated path stores rows four at a time"
    NQ = M // 4

    def quad_cost(q):
        rows = range(4 * q, 4 * q + 4)
        live = [m for m in rows if length[m] > 0]
        if not live:
            return 1
        bins = max(start[m] + length[m] for m in live) - min(start[m] for m in live)
        return sum(length[m] for m in rows) + bins + 1

    def partition(NP, extra):
        """optimal contiguous partition of the row quads into NP parts (minimise the heaviest part): DP over (quads, parts)"""
        cost = [quad_cost(q) + extra for q in range(NQ)]
        pre = [0]
        for c in cost:
            pre.append(pre[-1] + c)
        INF = float("inf")
        best = [[INF] * (NP + 1) for _ in range(NQ + 1)]
        cut = [[0] * (NP + 1) for _ in range(NQ + 1)]
        best[0][0] = 0
        for w in range(1, NP + 1):
            for q in range(NQ + 1):
                for j in range(q + 1):
                    v = max(best[j][w - 1], pre[q] - pre[j])
                    if v < best[q][w]:
                        best[q][w], cut[q][w] = v, j
        bounds = [NQ]
        q = NQ
        for w in range(NP, 0, -1):
            q = cut[q][w]
            bounds.append(q)
        bounds = bounds[::-1]
        return bounds, [sum(cost[bounds[w]:bounds[w + 1]]) for w in range(NP)]

    def emit_run(o, fn, NP, bounds, ret_range=False):
        o.append("  template <class C, class Emit4>")
        o.append(f"  static __device__ __forceinline__ {'int' if ret_range else 'void'} {fn}(int warp, const float* __restrict__ pr, Emit4&& emit4) {{")
        if ret_range:
            o.append("    int range = 0;  // first quad | (end quad << 8) of the part")
        o.append("    switch (warp) {")
        for w in range(NP):
            m0, m1 = 4 * bounds[w], 4 * bounds[w + 1]
            o.append(f"      case {w}: {{  // rows [{m0}, {m1})")
            if m1 > m0:
                rows = list(range(m0, m1))
                o.append("        float " + ", ".join(f"a{m} = 0.0f" for m in rows) + ";")
                live = [m for m in rows if length[m] > 0]
                if live:
                    k0 = min(start[m] for m in live)
                    k1 = max(start[m] + length[m] for m in live)
                    for k in range(k0, k1):
                        users = [m for m in live if start[m] <= k < start[m] + length[m] and fb[m][k] != 0.0]
                        if not users:
                            continue
                        stmts = " ".join(f"a{m} = fmaf(p, {_hexf(fb[m][k])}, a{m});" for m in users)
                        o.append(f"        {{ const float p = pr[2 * C::sig({k})]; {stmts} }}")
                for m in range(m0, m1, 4):
                    o.append(f"        emit4(std::integral_constant<int, {m}>{{}}, a{m}, a{m + 1}, a{m + 2}, a{m + 3});")
            if ret_range:
                o.append(f"        range = {bounds[w]} | ({bounds[w + 1]} << 8);")
            o.append("      } break;")
        o.append("      default: break;")
        o.append("    }")
        if ret_range:
            o.append("    return range;")
        o.append("  }")

    bounds, wcost = partition(NW, 0)
    o = []
    o.append(f"// {name}: sr={sr} n_fft={n_fft} n_mels={M} f_min={fmin} f_max={fmax or sr / 2} "
             f"norm={'slaney' if norm else 'none'} scale={'htk' if htk else 'slaney'}; {nnz} taps; "
             f"warp costs {wcost}")
    o.append(f"struct MelSpec_{name} {{")
    o.append(f"  static constexpr int M = {M}, F = {F}, NW = {NW}, NNZ = {nnz}, N_FFT = {n_fft};")
    o.append(f"  static constexpr const char* kName = \"{name}\";")
    o.append(f"  static constexpr int kStart[{M}] = {{{', '.join(map(str, start))}}};")
    o.append(f"  static constexpr int kLen[{M}] = {{{', '.join(map(str, length))}}};")
    o.append(f"  static constexpr unsigned kWBits[{max(nnz, 1)}] = {{{', '.join('0x%08xu' % b for b in wbits) or '0u'}}};")
    o.append("  // lane == frame: `pr` is this lane's power row inside the exchange buffer (bin k at float 2 * C::sig(k));")
    o.append("  // emit4(integral_constant<m0>, a0..a3) per row quad")
    emit_run(o, "run", NW, bounds)
    o.append("};")
    return "\n".join(o)


def generate(path=OUT):
    with tempfile.TemporaryDirectory() as tmp:
        lib = _tables_lib(tmp)
        parts = [
            "// GENERATED by csrc/gen_mel.py from b2a_mel_filters (csrc/tables.cu) — do not edit.",
            "// Straight-line mel projections for the named filterbanks; see gen_mel.py for the why and the how.",
            "#pragma once",
            "#include <type_traits>",
            "",
            "namespace b2a {",
            "namespace melgen {",
            "",
        ]
        for s in SPECS:
            parts.append(_emit_spec(lib, s))
            parts.append("")
        parts.append("#define B2A_MEL_SPECS(X) " + " ".join(f"X(MelSpec_{s[0]})" for s in SPECS))
        parts += ["", "}  // namespace melgen", "}  // namespace b2a", ""]
    text = "\n".join(parts)
    if os.path.exists(path) and open(path).read() == text:
        return path
    with open(path, "w") as f:
        f.write(text)
    return path


if __name__ == "__main__":
    print(generate())
