"""CPU: the C-ABI library loads, exports every symbol include/b200audio.h declares, and its host-side
tables / geometry agree with the oracle and the reference-over-shim fixtures.  No compute calls."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_build_and_load():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g

    g.build()
    from mlx_audio_plus_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH)
    assert _lib.lib.b2a_version() == 100


def test_every_declared_symbol_is_exported():
    from mlx_audio_plus_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "b200audio.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(b2a_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (b2a_[a-z0-9_]+)", nm))
    assert declared <= exported, declared - exported


def test_struct_sizes_match_header():
    """ctypes mirrors of the C structs: compile a tiny C program printing sizeof and compare."""
    from mlx_audio_plus_b200 import _lib

    src = ('#include <stdio.h>\n#include "b200audio.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n",'
           "sizeof(b2a_frontend_desc),sizeof(b2a_forward_args),sizeof(b2a_istft_desc),sizeof(b2a_inverse_args),"
           "sizeof(b2a_resample_args));return 0;}")
    import tempfile

    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "s.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(td, "s.c"), "-o", os.path.join(td, "s")])
        sizes = [int(v) for v in subprocess.check_output([os.path.join(td, "s")]).split()]
    assert sizes == [C.sizeof(_lib.FrontendDesc), C.sizeof(_lib.ForwardArgs), C.sizeof(_lib.IstftDesc), C.sizeof(_lib.InverseArgs),
                     C.sizeof(_lib.ResampleArgs)]


def test_windows_bit_exact_vs_oracle_and_fixture(golden):
    from mlx_audio_plus_b200 import dsp
    from oracle import dsp_oracle as O

    g = golden("tables")
    for kind in ("hanning", "hamming", "blackman", "bartlett"):
        for size in (16, 20, 21, 400, 401, 1024):
            for per in (False, True):
                w = np.asarray(getattr(dsp, kind)(size, per))
                np.testing.assert_array_equal(w, getattr(O, kind)(size, per))
                np.testing.assert_array_equal(w, g[f"win|{kind}|{size}|{int(per)}"])
    assert dsp.hanning(400) is dsp.hanning(400)  # lru-cached singleton like the reference
    with pytest.raises(ValueError):
        dsp.hanning(400).__setitem__(0, 1.0)  # read-only shared object


FB_CASES = {
    "whisper80": (16000, 400, 80, 0, None, "slaney", None),
    "whisper128": (16000, 400, 128, 0, None, "slaney", None),
    "parakeet80": (16000, 512, 80, 0, None, "per_feature", None),
    "vocos100": (24000, 1024, 100, 0, None, None, "htk"),
    "qwen3tts128": (24000, 1024, 128, 0.0, 12000.0, "slaney", "slaney"),
    "voxtral128": (16000, 400, 128, 0, 8000, "slaney", "slaney"),
    "funasr80": (16000, 400, 80, 0, None, "slaney", "htk"),
    "s3gen80": (24000, 1920, 80, 0, 8000, "slaney", "slaney"),
    "spark128": (16000, 1024, 128, 10, 8000, "slaney", "slaney"),
}


@pytest.mark.parametrize("name", sorted(FB_CASES))
def test_mel_filters_vs_fixture(golden, name):
    from mlx_audio_plus_b200 import dsp

    fb = np.asarray(dsp.mel_filters(*FB_CASES[name]))
    ref = golden("tables")[f"fb|{name}"]
    assert fb.shape == ref.shape and fb.dtype == np.float32
    # float32 libm exp/pow vs NumPy's differ in the last ulp of the band edges
    np.testing.assert_allclose(fb, ref, rtol=0, atol=1e-5 * np.abs(ref).max())
    assert ((fb != 0) == (ref != 0)).mean() > 0.9995
    assert (np.count_nonzero(fb, axis=0) <= 2).all()


def test_mel_filters_hashable_args_like_reference():
    from mlx_audio_plus_b200 import dsp

    with pytest.raises(TypeError):
        dsp.mel_filters(16000, 400, 80, f_max=np.zeros(2))  # unhashable -> lru_cache TypeError


@pytest.mark.parametrize("L,n_fft,hop,center,mode", [
    (4000, 400, 160, 1, 0), (4001, 400, 160, 1, 1), (4159, 512, 160, 1, 0), (999, 20, 5, 1, 0),
    (2048, 1024, 256, 0, 0), (201, 400, 160, 1, 0), (150, 400, 160, 1, 0), (200, 400, 160, 1, 0), (333, 16, 4, 1, 0)])
def test_geometry_and_index_map_bit_exact(L, n_fft, hop, center, mode):
    from mlx_audio_plus_b200 import _lib
    from oracle import dsp_oracle as O

    pm = "constant" if mode else "reflect"
    padded, T = C.c_int64(), C.c_int64()
    rc = _lib.lib.b2a_stft_geometry(L, n_fft, hop, center, mode, C.byref(padded), C.byref(T))
    try:
        idx = O.frame_indices(L, n_fft, hop, bool(center), pm)
    except ValueError:
        assert rc == _lib.ERR_TOO_SHORT
        return
    assert rc == 0 and T.value == idx.shape[0]
    rng = np.random.default_rng(0)
    for t in sorted(set([0, min(1, T.value - 1), T.value - 1] + list(rng.integers(0, T.value, 8)))):
        got = [_lib.lib.b2a_frame_source_index(L, n_fft, hop, center, mode, int(t), k) for k in range(n_fft)]
        np.testing.assert_array_equal(np.array(got), idx[int(t)])


def test_too_short_maps_to_value_error():
    from mlx_audio_plus_b200 import _lib

    rc = _lib.lib.b2a_stft_geometry(100, 400, 160, 0, 0, None, None)
    assert rc == _lib.ERR_TOO_SHORT
    with pytest.raises(ValueError, match="too short"):
        _lib.check(rc)


@pytest.mark.parametrize("T,n_fft,hop,center,length", [(468, 1024, 256, 1, -1), (200, 20, 5, 1, -1), (40, 64, 16, 1, 500),
                                                       (40, 64, 16, 0, -1), (3, 64, 16, 1, 10**6)])
def test_istft_geometry(T, n_fft, hop, center, length):
    from mlx_audio_plus_b200 import _lib
    from oracle import dsp_oracle as O

    ola, start, n = C.c_int64(), C.c_int64(), C.c_int64()
    assert _lib.lib.b2a_istft_geometry(T, n_fft, hop, center, length, C.byref(ola), C.byref(start), C.byref(n)) == 0
    spec = np.zeros((n_fft // 2 + 1, T), np.complex64)
    ref = O.istft(spec, hop, n_fft, np.ones(n_fft, np.float32), bool(center), None if length < 0 else length)
    assert n.value == ref.shape[0] and ola.value == (T - 1) * hop + n_fft
    if (T, n_fft) == (468, 1024):
        assert n.value == 119552  # codec/tests/test_vocos.py:61-73


def test_python_side_error_behaviour_without_gpu():
    from mlx_audio_plus_b200 import dsp

    with pytest.raises(ValueError, match="Unknown window function"):
        dsp.stft(np.zeros(1000, np.float32), 400, window="nope")
    with pytest.raises(ValueError, match="Invalid pad_mode"):
        dsp.stft(np.zeros(1000, np.float32), 400, pad_mode="edge")
    with pytest.raises(ValueError, match="Unknown window function"):
        dsp.istft(np.zeros((11, 5), np.complex64), 5, 20, window="nope")
    with pytest.raises(ValueError):
        dsp.istft(np.zeros((11, 50), np.complex64), 5, 21, dsp.hanning(21))  # window longer than irfft frame


def test_no_cpu_fallback():
    """Without a CUDA device the product path must fail loudly, never compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mlx_audio_plus_b200 import _lib, dsp

    with pytest.raises(_lib.B2AError, match="no CPU fallback"):
        dsp.stft(np.zeros(4000, np.float32), 400, 160)


def test_dsp_module_surface_and_isolation():
    """Mirrors mlx_audio/tests/test_dsp.py:7-53: import isolation + __all__ + utils re-export."""
    code = ("import sys; import mlx_audio_plus_b200.dsp as d; "
            "bad=[m for m in sys.modules if m.startswith('mlx_audio_plus_b200.') and m.split('.')[1] in ('stt','tts','codec','vad','sts')]; "
            "assert not bad, bad; print(sorted(d.__all__))")
    out = subprocess.check_output([sys.executable, "-c", code], cwd=ROOT, text=True)
    for name in ("hanning", "hamming", "blackman", "bartlett", "STR_TO_WINDOW_FN", "stft", "istft", "ISTFTCache", "mel_filters"):
        assert name in out
    from mlx_audio_plus_b200 import dsp, utils

    for name in ("hanning", "hamming", "blackman", "bartlett", "STR_TO_WINDOW_FN", "stft", "istft", "mel_filters"):
        assert getattr(utils, name) is getattr(dsp, name)
    c = dsp.ISTFTCache()
    c.get_norm_buffer(1920, 384, 1920, dsp.hamming(1920, False), 10)
    c.get_positions(10, 1920, 384)
    assert c.cache_info() == {"norm_buffers": 1, "position_indices": 1, "total_cached_items": 2}
    c.clear_cache()
    assert c.cache_info()["total_cached_items"] == 0
