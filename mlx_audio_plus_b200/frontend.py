"""Plan objects over the C ABI: FrontendPlan (stft / spectrogram / log-mel family) and IstftPlan.

A plan is the hashable tuple of the kwargs the reference keys its lru_caches on (dsp.py:33,223) plus the
window / filterbank bytes; plans are cached per CUDA device (one plan per GPU, usable from any stream).
"""
from __future__ import annotations

import ctypes as C
import threading
from collections import OrderedDict

import numpy as np

from . import _lib as L
from ._arrays import Ingested, emit, ingest


def _current_device_and_stream(device=None):
    import torch

    if not torch.cuda.is_available():
        raise L.B2AError("b200audio: no CUDA device — this library has no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    return dev, torch.cuda.current_stream(dev).cuda_stream


class _PlanBase:
    def __init__(self):
        self._h = C.c_void_p()
        # the *_host entry points stage through plan-owned device buffers and streams: one host call per plan at a time
        self._host_lock = threading.Lock()
        self._tls = threading.local()

    def close(self):
        if self._h:
            L.lib.b2a_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # best effort
        try:
            self.close()
        except Exception:
            pass

    @property
    def kernel_name(self) -> str:
        return (L.lib.b2a_plan_kernel_name(self._h) or b"").decode()


class FrontendPlan(_PlanBase):
    def __init__(self, *, n_fft, hop, window, center=True, pad_mode="reflect", preemph=0.0, drop_last=False,
                 spec_kind=L.SPEC_COMPLEX, spec_eps=0.0, filterbank=None, log_kind=L.LOG_NONE,
                 guard_kind=L.GUARD_NONE, guard_eps=0.0, clamp_kind=L.CLAMP_NONE, clamp_value=0.0,
                 affine_add=0.0, affine_div=0.0, norm_kind=L.NORM_NONE, norm_ddof=0, norm_eps=0.0,
                 out_layout=L.LAYOUT_TM, frame_len=0, frame_dc=False, frame_preemph=0.0, dither=0.0, out_dtype="float32"):
        super().__init__()
        window = np.ascontiguousarray(window, dtype=np.float32)
        if center and pad_mode not in ("reflect", "constant"):
            raise ValueError(f"Invalid pad_mode {pad_mode}")  # dsp.py:126
        d = L.FrontendDesc()
        d.n_fft, d.hop, d.center = int(n_fft), int(hop), int(bool(center))
        d.pad_mode = L.PAD_CONSTANT if pad_mode == "constant" else L.PAD_REFLECT
        d.window_len, d.preemph, d.drop_last = int(window.shape[0]), float(preemph), int(bool(drop_last))
        d.spec_kind, d.spec_eps = int(spec_kind), float(spec_eps)
        fb = None
        if filterbank is not None:
            fb = np.ascontiguousarray(filterbank, dtype=np.float32)
            if fb.ndim != 2 or fb.shape[1] != n_fft // 2 + 1:
                raise ValueError(f"filterbank shape {fb.shape} does not match n_fft={n_fft}")
            d.n_mels = int(fb.shape[0])
        d.log_kind, d.guard_kind, d.guard_eps = int(log_kind), int(guard_kind), float(guard_eps)
        d.clamp_kind, d.clamp_value = int(clamp_kind), float(clamp_value)
        d.affine_add, d.affine_div = float(affine_add), float(affine_div)
        d.norm_kind, d.norm_ddof, d.norm_eps = int(norm_kind), int(norm_ddof), float(norm_eps)
        d.out_layout = int(out_layout)
        d.frame_len, d.frame_dc = int(frame_len), int(bool(frame_dc))
        d.frame_preemph, d.dither = float(frame_preemph), float(dither)
        self.out_dtype = str(out_dtype).split(".")[-1]
        if self.out_dtype not in ("float32", "float16", "bfloat16"):
            raise ValueError(f"unsupported out_dtype {out_dtype}")
        d.out_dtype = {"float32": L.DTYPE_F32, "float16": L.DTYPE_F16, "bfloat16": L.DTYPE_BF16}[self.out_dtype]
        self.desc = d
        self.n_fft, self.hop, self.n_freqs = int(n_fft), int(hop), n_fft // 2 + 1
        self.n_out = d.n_mels if d.n_mels > 0 else self.n_freqs
        self.complex_out = d.n_mels == 0 and d.spec_kind == L.SPEC_COMPLEX
        L.check(L.lib.b2a_frontend_create(C.byref(d), window.ctypes.data_as(C.c_void_p),
                                          fb.ctypes.data_as(C.c_void_p) if fb is not None else None,
                                          C.byref(self._h)))

    # -- geometry ---------------------------------------------------------------------------------------
    def out_frames(self, length: int) -> int:
        n = C.c_int64()
        L.check(L.lib.b2a_frontend_out_frames(self._h, int(length), C.byref(n)))
        return n.value

    def out_shape(self, batch, frames):
        if self.desc.out_layout == L.LAYOUT_MT:
            return (batch, self.n_out, frames)
        return (batch, frames, self.n_out)

    def _args(self, audio_ptr, clip_stride, length, valid_length, batch, out_ptr, *, pad_value=0.0,
              sample_offset=0, frame_begin=0, frame_count=-1, clip_max=None, feat_sums=None):
        a = L.ForwardArgs()
        a.audio, a.clip_stride, a.length, a.valid_length = audio_ptr, clip_stride, length, valid_length
        a.pad_value, a.batch, a.sample_offset = pad_value, batch, sample_offset
        a.frame_begin, a.frame_count, a.out, a.out_clip_stride = frame_begin, frame_count, out_ptr, 0
        a.clip_max, a.feat_sums, a.workspace, a.workspace_bytes = clip_max, feat_sums, None, 0
        return a

    # -- execution --------------------------------------------------------------------------------------
    def run(self, ing: Ingested, *, length=None, pad_value=0.0, frame_count=None, seed=0):
        """ing.data: (B, L) float32 (host ndarray or torch CUDA tensor).  `length` > L adds virtual right
        padding with pad_value (whisper `padding`, parakeet pad_to).  `frame_count` keeps only the first frames
        (Kaldi snip_edges=False framing); `seed` keys the dither stream.  Returns (B, ...) in the same place."""
        x = ing.data
        B, Lx = int(x.shape[0]), int(x.shape[1])
        length = Lx if length is None else int(length)
        T = self.out_frames(length)
        if frame_count is not None:
            T = min(T, int(frame_count))
        shape = self.out_shape(B, T)
        if ing.on_device:
            import torch

            with torch.cuda.device(ing.device):
                out = torch.empty(shape, dtype=torch.complex64 if self.complex_out else getattr(torch, self.out_dtype),
                                  device=ing.device)
                st = torch.cuda.current_stream(ing.device).cuda_stream
                a = self._args(x.data_ptr(), Lx, length, Lx, B, out.data_ptr(), pad_value=pad_value, frame_count=T)
                a.seed = int(seed)
                if T > 0:
                    ws = self._call_workspace(a, ing.device)  # noqa: F841  (kept alive until the launch is enqueued)
                    L.check(L.lib.b2a_frontend_forward(self._h, C.byref(a), C.c_void_p(st)))
            return out
        _current_device_and_stream()  # fail loudly without a GPU
        if self.out_dtype == "bfloat16":
            raise TypeError("bfloat16 features need a torch CUDA input (NumPy has no bfloat16)")
        out = np.empty(shape, dtype=np.complex64 if self.complex_out else np.dtype(self.out_dtype))
        a = self._args(x.ctypes.data, Lx, length, Lx, B, out.ctypes.data, pad_value=pad_value, frame_count=T)
        a.seed = int(seed)
        if T > 0:
            with self._host_lock:
                L.check(L.lib.b2a_frontend_forward_host(self._h, C.byref(a)))
        return out

    def run_host_pcm16(self, pcm, out=None, *, length=None, pad_value=0.0):
        """Host int16 PCM (B, L) as a decoder delivers it (audio_io.py:258-262: float32 = int16 / 32768) -> features on the
        host, through ONE pipelined b2a_frontend_forward_host call: half the host-to-device bytes of the float32 entry, and
        with out_dtype float16 (what whisper.py:990-996 casts its segments to) half the device-to-host bytes too.
        `pcm` / `out` may be pinned (torch) buffers viewed as NumPy arrays."""
        pcm = np.ascontiguousarray(pcm)
        if pcm.dtype != np.int16 or pcm.ndim != 2:
            raise TypeError("run_host_pcm16 takes a (batch, samples) int16 array")
        _current_device_and_stream()
        B, Lx = int(pcm.shape[0]), int(pcm.shape[1])
        length = Lx if length is None else int(length)
        T = self.out_frames(length)
        shape = self.out_shape(B, T)
        if self.out_dtype == "bfloat16":
            raise TypeError("bfloat16 features need a torch CUDA input (NumPy has no bfloat16)")
        if out is None:
            out = np.empty(shape, dtype=np.dtype(self.out_dtype))
        a = self._args(pcm.ctypes.data, Lx, length, Lx, B, out.ctypes.data, pad_value=pad_value, frame_count=T)
        a.audio_kind = L.PCM_I16
        if T > 0:
            with self._host_lock:
                L.check(L.lib.b2a_frontend_forward_host(self._h, C.byref(a)))
        return out

    def _call_workspace(self, a, device):
        """Per-call device scratch (clip maxima, per-feature sums, per-tile minima) from torch's caching allocator on the
        caller's current stream: the plan object is shared by every caller on the device (cached_plan), and with its own
        scratch two streams / threads running a clamping or normalising front-end would race on the statistics.  The
        allocator recycles the block in stream order, so it may be dropped as soon as the launch is enqueued."""
        if self.desc.clamp_kind == L.CLAMP_NONE and self.desc.norm_kind == L.NORM_NONE and not a.clip_max and not a.feat_sums:
            return None
        import torch

        n = int(L.lib.b2a_frontend_call_workspace_bytes(self._h, C.byref(a)))
        if n <= 0:
            return None
        ws = torch.empty(n, dtype=torch.uint8, device=device)
        a.workspace, a.workspace_bytes = ws.data_ptr(), n
        return ws

    # -- split form for frame-range sharding ------------------------------------------------
    def stats_tensors(self, batch, device):
        """(clip_max float32 [B], feat_sums float64 [B, n_out, 2]) — the buffers ranks all-reduce."""
        import torch

        return (torch.full((batch,), float("-inf"), dtype=torch.float32, device=device),
                torch.zeros((batch, self.n_out, 2), dtype=torch.float64, device=device))

    def partial(self, x_cuda, out, clip_max, feat_sums, *, length, sample_offset, frame_begin, frame_count,
                valid_length=None, pad_value=0.0):
        """Un-clamped / un-normalised features of frames [frame_begin, +frame_count) of a signal of GLOBAL
        `length`, from a slice whose first element is global sample `sample_offset`.  Statistics land in
        clip_max / feat_sums for the caller to reduce across ranks."""
        import torch

        B = int(x_cuda.shape[0])
        valid = length if valid_length is None else valid_length
        a = self._args(x_cuda.data_ptr(), int(x_cuda.shape[1]), int(length), int(valid), B, out.data_ptr(),
                       pad_value=pad_value, sample_offset=int(sample_offset), frame_begin=int(frame_begin),
                       frame_count=int(frame_count), clip_max=clip_max.data_ptr(), feat_sums=feat_sums.data_ptr())
        st = torch.cuda.current_stream(x_cuda.device).cuda_stream
        with torch.cuda.device(x_cuda.device):
            ws = self._call_workspace(a, x_cuda.device)
            L.check(L.lib.b2a_frontend_partial(self._h, C.byref(a), C.c_void_p(st)))
        # the call's context (arguments + its scratch, which finalize() reads the per-tile minima from): per thread, never
        # on the shared plan object
        self._tls.last_partial = (a, ws)
        return out

    def finalize(self, out, clip_max, feat_sums, *, global_frames):
        """Clamp / normalise `out` in place with the (all-reduced) statistics of this thread's last partial() call."""
        import torch

        a, _ws = self._tls.last_partial
        a.out, a.clip_max, a.feat_sums = out.data_ptr(), clip_max.data_ptr(), feat_sums.data_ptr()
        st = torch.cuda.current_stream(out.device).cuda_stream
        L.check(L.lib.b2a_frontend_finalize(self._h, C.byref(a), int(global_frames), C.c_void_p(st)))
        return out

    def dump_frames(self, x_cuda, apply_window=False, length=None):
        """Parity hook (bit-exact framing test): (B, L) torch CUDA float32 -> (B, T, n_fft) frames.
        Only meaningful on plans created with drop_last=False."""
        import torch

        B, Lx = x_cuda.shape
        length = Lx if length is None else int(length)
        T = self.out_frames(length)
        out = torch.empty((B, T, self.n_fft), dtype=torch.float32, device=x_cuda.device)
        a = self._args(x_cuda.data_ptr(), Lx, length, Lx, B, out.data_ptr(), frame_count=T)
        st = torch.cuda.current_stream(x_cuda.device).cuda_stream
        L.check(L.lib.b2a_frontend_dump_frames(self._h, C.byref(a), int(apply_window), C.c_void_p(st)))
        return out


class IstftPlan(_PlanBase):
    def __init__(self, *, n_fft, hop, window, center=True, normalized=False, div_clamp=False, trim_tail=True,
                 div_eps=0.0, polar=False, mag_clip_max=0.0, mag_clip_min_zero=False, mag_log=False):
        super().__init__()
        window = np.ascontiguousarray(window, dtype=np.float32)
        d = L.IstftDesc()
        d.n_fft, d.hop, d.window_len, d.center = int(n_fft), int(hop), int(window.shape[0]), int(bool(center))
        d.norm_kind = L.ISTFT_NORM_WINDOW_SQ if normalized else L.ISTFT_NORM_WINDOW
        d.div_kind = L.ISTFT_DIV_CLAMP if div_clamp else L.ISTFT_DIV_WHERE
        d.trim_tail = int(bool(trim_tail))
        d.div_eps = float(div_eps)
        d.input_form = L.ISTFT_INPUT_POLAR if polar else L.ISTFT_INPUT_COMPLEX
        d.mag_clip_max, d.mag_clip_min_zero = float(mag_clip_max), int(bool(mag_clip_min_zero))
        d.mag_log = int(bool(mag_log))
        self.desc = d
        self.n_fft, self.hop, self.n_freqs = int(n_fft), int(hop), n_fft // 2 + 1
        L.check(L.lib.b2a_istft_create(C.byref(d), window.ctypes.data_as(C.c_void_p), C.byref(self._h)))

    def out_len(self, num_frames, length=None) -> int:
        n = C.c_int64()
        L.check(L.lib.b2a_istft_out_len(self._h, int(num_frames), -1 if length is None else int(length), C.byref(n)))
        return n.value

    def run(self, ing: Ingested, imag: Ingested = None, *, length=None, clip_stride=0):
        """ing.data: (B, F, T) complex64 — or float32 real plane with `imag` the imaginary plane.  `clip_stride` (elements
        between clips; 0 = dense F * T) lets the two planes be halves of one (B, 2F, T) buffer."""
        x = ing.data
        B, F, T = (int(s) for s in x.shape)
        if F != self.n_freqs:
            raise ValueError(f"istft: {F} frequency bins do not match n_fft={self.n_fft} (needs {self.n_freqs})")
        n_out = self.out_len(T, length)
        a = L.InverseArgs()
        a.clip_stride, a.num_frames, a.batch = int(clip_stride), T, B
        a.length, a.out_clip_stride = (-1 if length is None else int(length)), 0
        if ing.on_device:
            import torch

            with torch.cuda.device(ing.device):
                out = torch.empty((B, n_out), dtype=torch.float32, device=ing.device)
                a.spec, a.spec_imag, a.out = x.data_ptr(), (imag.data.data_ptr() if imag is not None else None), out.data_ptr()
                st = torch.cuda.current_stream(ing.device).cuda_stream
                if n_out > 0:
                    L.check(L.lib.b2a_istft_inverse(self._h, C.byref(a), C.c_void_p(st)))
            return out
        _current_device_and_stream()
        out = np.empty((B, n_out), dtype=np.float32)
        a.spec, a.spec_imag, a.out = x.ctypes.data, (imag.data.ctypes.data if imag is not None else None), out.ctypes.data
        if n_out > 0:
            with self._host_lock:
                L.check(L.lib.b2a_istft_inverse_host(self._h, C.byref(a)))
        return out


# ---- per-device plan cache ------------------------------------------------------------------------------
_CACHE = OrderedDict()
_CACHE_LOCK = threading.Lock()
_CACHE_MAX = 64


def _device_index(ing: Ingested = None) -> int:
    import torch

    if ing is not None and ing.on_device:
        return ing.device.index if ing.device.index is not None else torch.cuda.current_device()
    if not torch.cuda.is_available():
        raise L.B2AError("b200audio: no CUDA device — this library has no CPU fallback")
    return torch.cuda.current_device()


_FP_CACHE = OrderedDict()  # (data pointer, shape, strides, dtype) of a READ-ONLY array -> (owner kept alive, fingerprint)
_FP_MAX = 64


def _fingerprint(arr):
    """Content key of a window / filterbank array.  The wrappers pass the same lru-cached, read-only tables (or views of
    them) on every call — dsp.hanning / dsp.mel_filters cache like the reference's do — so the ~100 KB filterbank is
    hashed once per table, not once per call.  Writable arrays may change under the same address and are hashed every time."""
    if arr is None:
        return None
    if isinstance(arr, np.ndarray) and not arr.flags.writeable:
        k = (arr.__array_interface__["data"][0], arr.shape, arr.strides, arr.dtype.str)
        with _CACHE_LOCK:
            ent = _FP_CACHE.get(k)
            if ent is not None:
                _FP_CACHE.move_to_end(k)
                return ent[1]
    else:
        k = None
    a = np.ascontiguousarray(arr, dtype=np.float32)
    fp = (a.shape, hash(a.tobytes()), float(a.sum(dtype=np.float64)))
    if k is not None:
        with _CACHE_LOCK:
            _FP_CACHE[k] = (arr, fp)  # holding the array keeps its address from being reused while the entry lives
            while len(_FP_CACHE) > _FP_MAX:
                _FP_CACHE.popitem(last=False)
    return fp


def cached_plan(cls, dev_index: int, window: np.ndarray, filterbank=None, **kw):
    key = (cls.__name__, dev_index, tuple(sorted(kw.items())), _fingerprint(window), _fingerprint(filterbank))
    with _CACHE_LOCK:
        plan = _CACHE.get(key)
        if plan is not None:
            _CACHE.move_to_end(key)
            return plan
    import torch

    window = np.ascontiguousarray(window, dtype=np.float32)
    with torch.cuda.device(dev_index):
        plan = cls(window=window, filterbank=filterbank, **kw) if filterbank is not None or cls is FrontendPlan \
            else cls(window=window, **kw)
    with _CACHE_LOCK:
        _CACHE[key] = plan
        while len(_CACHE) > _CACHE_MAX:
            _CACHE.popitem(last=False)
    return plan


def clear_plan_cache():
    with _CACHE_LOCK:
        _CACHE.clear()
