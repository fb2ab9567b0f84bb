"""ctypes binding of libvoxtral_b200.so, shaped like the reference crate's API (see __init__).

Error behaviour mirrors the reference: loader/IO problems surface as exceptions with the
reference's message texts where it has them (``anyhow`` contexts in reader.rs / loader.rs);
shape panics (op.rs:92-100) become ``VoxtralError`` instead of process aborts.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libvoxtral_b200.so"


class VoxtralError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[vox {code}] {msg}")
        self.code = code
        self.msg = msg


def lib_path() -> str:
    # VOX_LIB_PATH: load another build of the same library (A/B runs of two kernel versions on one GPU box)
    return os.environ.get("VOX_LIB_PATH") or os.path.join(_HERE, _LIB_NAME)


_lib = None


class _PadConfig(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("n_left_pad_tokens", C.c_uint32), ("frame_rate", C.c_float),
                ("extra_right_pad_tokens", C.c_uint32)]


class _Chunk(C.Structure):
    _fields_ = [("start_sample", C.c_size_t), ("end_sample", C.c_size_t), ("index", C.c_size_t),
                ("is_last", C.c_int32)]


class _ModelInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_mels", "enc_dim", "enc_layers", "enc_heads", "enc_head_dim", "enc_ffn", "enc_window",
        "dec_dim", "dec_layers", "dec_heads", "dec_kv_heads", "dec_head_dim", "dec_ffn", "dec_window",
        "vocab", "t_cond_dim", "reshape_factor", "prefix_len")] + [
        ("q4_bytes", C.c_uint64), ("device_bytes", C.c_uint64), ("decode_step_bytes", C.c_uint64)]


class _Timings(C.Structure):
    _fields_ = [("preprocess_ms", C.c_float), ("encode_ms", C.c_float), ("decode_ms", C.c_float),
                ("total_ms", C.c_float), ("prefill_ms", C.c_float), ("decode_tokens", C.c_int32), ("seq_len", C.c_int32)]


@dataclass
class Timings:
    preprocess_ms: float = 0.0
    encode_ms: float = 0.0
    decode_ms: float = 0.0
    total_ms: float = 0.0
    prefill_ms: float = 0.0
    decode_tokens: int = 0
    seq_len: int = 0


# name -> (restype, argtypes); every declaration in include/voxtral.h appears here
_P = C.c_void_p
_SIGS = {
    "vox_last_error": (C.c_char_p, []),
    "vox_version": (C.c_int32, []),
    "vox_device_count": (C.c_int32, []),
    "vox_gguf_open": (C.c_int32, [C.c_char_p, C.POINTER(_P)]),
    "vox_gguf_open_shards": (C.c_int32, [C.POINTER(_P), C.POINTER(C.c_size_t), C.c_size_t, C.POINTER(_P)]),
    "vox_gguf_version": (C.c_int32, [_P, C.POINTER(C.c_uint32)]),
    "vox_gguf_tensor_count": (C.c_int32, [_P, C.POINTER(C.c_uint64)]),
    "vox_gguf_tensor_name": (C.c_int32, [_P, C.c_uint64, C.POINTER(C.c_char_p)]),
    "vox_gguf_tensor_info": (C.c_int32, [_P, C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                         C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "vox_gguf_tensor_data": (C.c_int32, [_P, C.c_char_p, _P, C.c_size_t]),
    "vox_gguf_close": (None, [_P]),
    "vox_peak_normalize": (C.c_int32, [_P, C.c_size_t, C.c_float]),
    "vox_pad_config_default": (None, [C.POINTER(_PadConfig)]),
    "vox_pad_audio_len": (C.c_size_t, [C.c_size_t, C.POINTER(_PadConfig)]),
    "vox_pad_audio": (C.c_int32, [_P, C.c_size_t, C.POINTER(_PadConfig), _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "vox_stream_progress": (C.c_int32, [C.c_size_t, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]),
    "vox_chunk_plan": (C.c_int32, [C.c_size_t, C.c_size_t, C.c_size_t, C.POINTER(_Chunk), C.c_size_t,
                                   C.POINTER(C.c_size_t)]),
    "vox_time_embedding": (C.c_int32, [C.c_float, C.c_int32, _P]),
    "vox_mel_create": (C.c_int32, [C.c_int32, C.POINTER(_P)]),
    "vox_mel_num_frames": (C.c_size_t, [C.c_size_t]),
    "vox_mel_compute_log": (C.c_int32, [_P, _P, C.c_size_t, _P, C.c_size_t]),
    "vox_mel_compute_log_dev": (C.c_int32, [_P, _P, C.c_size_t, _P, C.c_int32, _P]),
    "vox_mel_filterbank": (C.c_int32, [_P, _P]),
    "vox_mel_window": (C.c_int32, [_P, _P]),
    "vox_mel_free": (None, [_P]),
    "vox_q4_tensor_create": (C.c_int32, [_P, C.c_size_t, C.c_int64, C.c_int64, C.c_int32, C.POINTER(_P)]),
    "vox_q4_tensor_shape": (C.c_int32, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "vox_q4_tensor_dequantize": (C.c_int32, [_P, _P]),
    "vox_q4_matmul": (C.c_int32, [_P, _P, _P, C.c_int32, C.c_int32, _P, _P]),
    "vox_q4_matmul_host": (C.c_int32, [_P, _P, _P, C.c_int32, C.c_int32, _P]),
    "vox_q4_tensor_free": (None, [_P]),
    "vox_q4_set_matvec_mode": (C.c_int32, [C.c_int32]),
    "vox_dev_malloc": (C.c_int32, [C.c_int32, C.c_size_t, C.POINTER(_P)]),
    "vox_dev_free": (C.c_int32, [C.c_int32, _P]),
    "vox_dev_upload": (C.c_int32, [C.c_int32, _P, _P, C.c_size_t]),
    "vox_dev_download": (C.c_int32, [C.c_int32, _P, _P, C.c_size_t]),
    "vox_dev_sync": (C.c_int32, [C.c_int32]),
    "vox_profiler_start": (C.c_int32, []),
    "vox_profiler_stop": (C.c_int32, []),
    "vox_host_alloc_pinned": (C.c_int32, [C.c_size_t, C.POINTER(_P)]),
    "vox_host_free_pinned": (C.c_int32, [_P]),
    "vox_q4_matmul_bench": (C.c_int32, [C.POINTER(_P), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.POINTER(C.c_float)]),
    "vox_model_load_gguf": (C.c_int32, [C.c_char_p, C.c_int32, C.POINTER(_P)]),
    "vox_model_load_gguf_handle": (C.c_int32, [_P, C.c_int32, C.POINTER(_P)]),
    "vox_model_get_info": (C.c_int32, [_P, C.POINTER(_ModelInfo)]),
    "vox_model_free": (None, [_P]),
    "vox_session_create": (C.c_int32, [_P, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "vox_session_set_delay": (C.c_int32, [_P, C.c_float]),
    "vox_encode_audio": (C.c_int32, [_P, _P, C.c_int32, C.c_int32, _P, C.c_size_t, C.POINTER(C.c_int32)]),
    "vox_transcribe_streaming": (C.c_int32, [_P, _P, C.c_int32, C.c_int32, _P, C.c_size_t, C.POINTER(C.c_int32),
                                             C.POINTER(_Timings)]),
    "vox_transcribe_pcm": (C.c_int32, [_P, _P, C.c_int32, C.c_size_t, C.c_int32, _P, C.c_size_t,
                                       C.POINTER(C.c_int32), C.POINTER(_Timings)]),
    "vox_transcribe_pcm_dev": (C.c_int32, [_P, _P, C.c_int32, C.c_size_t, _P, C.c_size_t, C.POINTER(C.c_int32),
                                           C.POINTER(_Timings)]),
    "vox_generate_step_with_cache": (C.c_int32, [_P, _P, C.c_int32, C.c_int32, _P, C.c_size_t]),
    "vox_forward_streaming": (C.c_int32, [_P, _P, C.c_int32, C.c_int32, _P, C.c_int32, _P, C.c_size_t]),
    "vox_prefill": (C.c_int32, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "vox_decode_step": (C.c_int32, [_P, _P, C.c_int32, C.c_int32, _P]),
    "vox_session_cache_len": (C.c_int32, [_P, C.POINTER(C.c_int32)]),
    "vox_session_reset": (C.c_int32, [_P]),
    "vox_session_debug_read": (C.c_int32, [_P, C.c_char_p, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "vox_session_launch_count": (C.c_int32, [_P, C.POINTER(C.c_uint64)]),
    "vox_session_free": (None, [_P]),
    "vox_stream_pool_create": (C.c_int32, [_P, C.c_int32, C.c_float, C.POINTER(_P)]),
    "vox_stream_open": (C.c_int32, [_P, C.POINTER(C.c_int32)]),
    "vox_stream_push_pcm": (C.c_int32, [_P, C.c_int32, _P, C.c_size_t]),
    "vox_stream_finish": (C.c_int32, [_P, C.c_int32]),
    "vox_stream_tick": (C.c_int32, [_P, _P]),
    "vox_stream_poll_ids": (C.c_int32, [_P, C.c_int32, _P, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_int32)]),
    "vox_stream_audio_embeds": (C.c_int32, [_P, C.c_int32, _P, C.c_size_t, C.POINTER(C.c_int32)]),
    "vox_stream_encode_chunk": (C.c_int32, [_P, C.c_int32, _P, C.c_int32, _P, C.c_size_t, C.POINTER(C.c_int32)]),
    "vox_stream_close": (C.c_int32, [_P, C.c_int32]),
    "vox_stream_pool_free": (None, [_P]),
    "vox_tokenizer_from_file": (C.c_int32, [C.c_char_p, C.POINTER(_P)]),
    "vox_tokenizer_from_json": (C.c_int32, [C.c_char_p, C.c_size_t, C.POINTER(_P)]),
    "vox_tokenizer_decode": (C.c_int32, [_P, _P, C.c_size_t, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "vox_tokenizer_decode_token": (C.c_int32, [_P, C.c_uint32, _P, C.c_size_t, C.POINTER(C.c_size_t),
                                               C.POINTER(C.c_int32)]),
    "vox_tokenizer_vocab_size": (C.c_int32, [_P, C.POINTER(C.c_size_t)]),
    "vox_tokenizer_free": (None, [_P]),
}


def lib():
    """Load libvoxtral_b200.so (fails loudly if it was not built: there is no fallback)."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise VoxtralError(-1, f"{path} not found -- run `python -m voxtral_mini_realtime_rs_b200.build` "
                                   "(or __graft_entry__.build()); there is no CPU/PyTorch fallback")
        l = C.CDLL(path)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def _check(code: int):
    if code != 0:
        raise VoxtralError(code, lib().vox_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    return int(lib().vox_device_count())


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------------------ GGUF
class GgufReader:
    """GgufReader (reader.rs:88-223).  `open(path)`, `from_bytes(b)`, `from_shards([b0,b1,..])`."""

    def __init__(self, handle, keep=None):
        self._h = handle
        self._keep = keep

    @staticmethod
    def open(path: str) -> "GgufReader":
        h = _P()
        _check(lib().vox_gguf_open(os.fsencode(path), C.byref(h)))
        return GgufReader(h)

    @staticmethod
    def from_shards(shards) -> "GgufReader":
        bufs = [np.frombuffer(s, dtype=np.uint8) for s in shards]
        n = len(bufs)
        ptrs = (_P * n)(*[b.ctypes.data for b in bufs])
        lens = (C.c_size_t * n)(*[b.size for b in bufs])
        h = _P()
        _check(lib().vox_gguf_open_shards(ptrs, lens, n, C.byref(h)))
        return GgufReader(h, keep=(bufs, shards))

    @staticmethod
    def from_bytes(data) -> "GgufReader":
        return GgufReader.from_shards([data])

    def version(self) -> int:
        v = C.c_uint32()
        _check(lib().vox_gguf_version(self._h, C.byref(v)))
        return v.value

    def tensor_count(self) -> int:
        v = C.c_uint64()
        _check(lib().vox_gguf_tensor_count(self._h, C.byref(v)))
        return v.value

    def tensor_names(self):
        out = []
        for i in range(self.tensor_count()):
            s = C.c_char_p()
            _check(lib().vox_gguf_tensor_name(self._h, i, C.byref(s)))
            out.append(s.value.decode())
        return out

    def tensor_info(self, name: str):
        """-> dict(shape=<GGUF-order dims>, dtype=<code>, nbytes=...) or None (reader.rs:201)."""
        dt, nd, nb = C.c_uint32(), C.c_uint32(), C.c_uint64()
        dims = (C.c_uint64 * 4)()
        code = lib().vox_gguf_tensor_info(self._h, name.encode(), C.byref(dt), C.byref(nd), dims, C.byref(nb))
        if code == 3:
            return None
        _check(code)
        return dict(shape=tuple(dims[i] for i in range(nd.value)), dtype=dt.value, nbytes=nb.value)

    def tensor_data(self, name: str) -> np.ndarray:
        info = self.tensor_info(name)
        if info is None:
            raise VoxtralError(3, f"Tensor '{name}' not found in GGUF")
        buf = np.empty(info["nbytes"], np.uint8)
        _check(lib().vox_gguf_tensor_data(self._h, name.encode(), _ptr(buf), buf.size))
        return buf

    def close(self):
        if self._h:
            lib().vox_gguf_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------ audio plumbing
def peak_normalize(samples, target_peak: float = 0.95) -> np.ndarray:
    """AudioBuffer::peak_normalize (io.rs:59-68); returns a new array."""
    s = _f32(samples).copy()
    _check(lib().vox_peak_normalize(_ptr(s), s.size, target_peak))
    return s


class PadConfig:
    """PadConfig (pad.rs:20-46)."""

    def __init__(self, sample_rate=16000, n_left_pad_tokens=76, frame_rate=12.5, extra_right_pad_tokens=17):
        self.c = _PadConfig(sample_rate, n_left_pad_tokens, frame_rate, extra_right_pad_tokens)

    @staticmethod
    def voxtral() -> "PadConfig":
        return PadConfig()

    def samples_per_token(self) -> int:
        return int(np.float32(self.c.sample_rate) / np.float32(self.c.frame_rate))

    def left_pad_samples(self) -> int:
        return self.c.n_left_pad_tokens * self.samples_per_token()


def pad_audio(samples, config: PadConfig | None = None) -> np.ndarray:
    """pad_audio (pad.rs:89-103)."""
    s = _f32(samples)
    cfg = C.byref(config.c) if config else None
    n = lib().vox_pad_audio_len(s.size, cfg)
    out = np.empty(n, np.float32)
    ln = C.c_size_t()
    _check(lib().vox_pad_audio(_ptr(s), s.size, cfg, _ptr(out), out.size, C.byref(ln)))
    return out[:ln.value]


def chunk_audio(n_samples: int, max_mel_frames: int = 1500, overlap_frames: int = 0):
    """chunk_audio (chunk.rs:159-161) as a plan: [(start, end, index, is_last)]."""
    cnt = C.c_size_t()
    _check(lib().vox_chunk_plan(n_samples, max_mel_frames, overlap_frames, None, 0, C.byref(cnt)))
    arr = (_Chunk * max(cnt.value, 1))()
    _check(lib().vox_chunk_plan(n_samples, max_mel_frames, overlap_frames, arr, cnt.value, C.byref(cnt)))
    return [(arr[i].start_sample, arr[i].end_sample, arr[i].index, bool(arr[i].is_last)) for i in range(cnt.value)]


def stream_progress(n_samples: int, ended: bool = False, reshape_factor: int = 4, prefix_len: int = 38):
    """Final outputs per stage once `n_samples` padded samples are known: (mel frames, conv1 frames, encoder
    frames, audio embeddings, emit-able token ids) -- the bookkeeping of a streaming session."""
    out = (C.c_int64 * 5)()
    _check(lib().vox_stream_progress(n_samples, 1 if ended else 0, reshape_factor, prefix_len, out))
    return tuple(int(v) for v in out)


def needs_chunking(n_samples: int, max_mel_frames: int = 1500) -> bool:
    return n_samples > max_mel_frames * 160


class TimeEmbedding:
    """TimeEmbedding (time_embedding.rs:12-71)."""

    def __init__(self, dim: int):
        self.dim = dim

    def embed(self, t: float) -> np.ndarray:
        out = np.empty(self.dim, np.float32)
        _check(lib().vox_time_embedding(t, self.dim, _ptr(out)))
        return out.reshape(1, 1, self.dim)


# ------------------------------------------------------------------------------ mel
class MelSpectrogram:
    """MelSpectrogram::voxtral() (mel.rs:73-182) on the GPU."""

    def __init__(self, device: int = 0):
        self._h = _P()
        self.device = device
        _check(lib().vox_mel_create(device, C.byref(self._h)))

    @staticmethod
    def voxtral(device: int = 0) -> "MelSpectrogram":
        return MelSpectrogram(device)

    @staticmethod
    def num_frames(num_samples: int) -> int:
        return int(lib().vox_mel_num_frames(num_samples))

    def compute_log(self, samples) -> np.ndarray:
        """-> float32 [n_frames, 128] (mel.rs:128-165)."""
        s = _f32(samples)
        fr = self.num_frames(s.size)
        out = np.empty((fr, 128), np.float32)
        _check(lib().vox_mel_compute_log(self._h, _ptr(s), s.size, _ptr(out), out.size))
        return out

    def mel_basis(self) -> np.ndarray:
        out = np.empty((128, 201), np.float32)
        _check(lib().vox_mel_filterbank(self._h, _ptr(out)))
        return out

    def window(self) -> np.ndarray:
        out = np.empty(400, np.float32)
        _check(lib().vox_mel_window(self._h, _ptr(out)))
        return out

    def __del__(self):
        try:
            if self._h:
                lib().vox_mel_free(self._h)
                self._h = None
        except Exception:
            pass


# ------------------------------------------------------------------------------ Q4 operator
class DeviceBuffer:
    """Raw device allocation for driving the *_dev entry points without torch."""

    def __init__(self, nbytes: int, device: int = 0):
        self.device, self.nbytes = device, nbytes
        self.ptr = _P()
        _check(lib().vox_dev_malloc(device, nbytes, C.byref(self.ptr)))

    @staticmethod
    def from_numpy(a: np.ndarray, device: int = 0) -> "DeviceBuffer":
        a = np.ascontiguousarray(a)
        b = DeviceBuffer(a.nbytes, device)
        _check(lib().vox_dev_upload(device, b.ptr, _ptr(a), a.nbytes))
        return b

    def to_numpy(self, dtype, shape) -> np.ndarray:
        out = np.empty(shape, dtype)
        _check(lib().vox_dev_download(self.device, _ptr(out), self.ptr, out.nbytes))
        return out

    def free(self):
        if self.ptr:
            lib().vox_dev_free(self.device, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PinnedArray:
    """float32/int32 numpy view over page-locked host memory (cudaHostAlloc)."""

    def __init__(self, shape, dtype=np.float32):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = _P()
        _check(lib().vox_host_alloc_pinned(max(self.nbytes, 16), C.byref(self.ptr)))
        buf = (C.c_char * max(self.nbytes, 16)).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().vox_host_free_pinned(self.ptr)
            self.ptr = None


class Q4Tensor:
    """Q4Tensor (tensor.rs:16-113): raw Q4_0 blocks -> HBM (repacked at upload)."""

    def __init__(self, handle, shape, device):
        self._h, self._shape, self.device = handle, tuple(shape), device

    @staticmethod
    def from_q4_bytes(raw_bytes, shape, device: int = 0) -> "Q4Tensor":
        raw = np.ascontiguousarray(np.frombuffer(raw_bytes, dtype=np.uint8) if not isinstance(raw_bytes, np.ndarray)
                                   else raw_bytes.astype(np.uint8, copy=False))
        n, k = int(shape[0]), int(shape[1])
        h = _P()
        _check(lib().vox_q4_tensor_create(_ptr(raw), raw.size, n, k, device, C.byref(h)))
        return Q4Tensor(h, (n, k), device)

    def shape(self):
        return self._shape

    def num_blocks(self) -> int:
        return self._shape[0] * self._shape[1] // 32

    def dequantize(self) -> np.ndarray:
        out = np.empty(self._shape, np.float32)
        _check(lib().vox_q4_tensor_dequantize(self._h, _ptr(out)))
        return out

    def __del__(self):
        try:
            if self._h:
                lib().vox_q4_tensor_free(self._h)
                self._h = None
        except Exception:
            pass


def q4_matmul(x, weights: Q4Tensor, bias=None) -> np.ndarray:
    """q4_matmul (op.rs:86-137): x [B,M,K] host f32 -> [B,M,N] host f32 (upload, kernel, download)."""
    x = _f32(x)
    if x.ndim != 3:
        raise VoxtralError(1, "Input must be 3D [B, M, K]")
    b, m, k = x.shape
    n, wk = weights.shape()
    if k != wk:
        raise VoxtralError(1, f"K dimension mismatch: input has {k}, weights have {wk}")
    y = np.empty((b, m, n), np.float32)
    bp = None
    if bias is not None:
        bias = _f32(bias)
        if bias.size != n:
            raise VoxtralError(1, f"bias has {bias.size} elements, expected {n}")
        bp = _ptr(bias)
    _check(lib().vox_q4_matmul_host(weights._h, _ptr(x), _ptr(y), b, m, bp))
    return y


class Q4Linear:
    """Q4Linear (linear.rs:17-40)."""

    def __init__(self, weights: Q4Tensor, bias=None):
        self.weights, self.bias = weights, bias

    def forward(self, x) -> np.ndarray:
        return q4_matmul(x, self.weights, self.bias)


def q4_matmul_bench(weights, m: int, iters: int = 200, warmup: int = 20) -> float:
    """Average ms per launch over `iters` launches rotating over `weights` (defeats L2)."""
    n = len(weights)
    arr = (_P * n)(*[w._h for w in weights])
    ms = C.c_float()
    _check(lib().vox_q4_matmul_bench(arr, n, m, iters, warmup, C.byref(ms)))
    return ms.value


# ------------------------------------------------------------------------------ model
class Q4VoxtralModel:
    """Q4VoxtralModel (model.rs:759-989) + its session state (LayerCaches, workspace, stream)."""

    def __init__(self, model_handle, device: int, max_batch: int = 1, max_mel_frames: int = 3000):
        self._m = model_handle
        self.device = device
        info = _ModelInfo()
        _check(lib().vox_model_get_info(self._m, C.byref(info)))
        self.info = {f[0]: getattr(info, f[0]) for f in _ModelInfo._fields_}
        self._s = _P()
        self.max_batch, self.max_mel_frames = max_batch, max_mel_frames
        _check(lib().vox_session_create(self._m, max_batch, max_mel_frames, C.byref(self._s)))

    def set_delay(self, delay_tokens: float):
        _check(lib().vox_session_set_delay(self._s, delay_tokens))

    def _mel3(self, mel):
        mel = _f32(mel)
        if mel.ndim == 2:
            mel = mel[None]
        if mel.ndim != 3 or mel.shape[1] != self.info["n_mels"]:
            raise VoxtralError(1, f"mel must be [B,{self.info['n_mels']},T], got {mel.shape}")
        return mel

    def encode_audio(self, mel) -> np.ndarray:
        """mel [B,128,T] -> audio embeds [B, T/16, dec_dim] (model.rs:783-788)."""
        mel = self._mel3(mel)
        b, _, t = mel.shape
        t1 = (t + 2 - 3) // 2 + 1
        s = (t1 + 2 - 3) // 2 + 1
        s4 = s // self.info["reshape_factor"]
        out = np.empty((b, s4, self.info["dec_dim"]), np.float32)
        sl = C.c_int32()
        _check(lib().vox_encode_audio(self._s, _ptr(mel), b, t, _ptr(out), out.size, C.byref(sl)))
        assert sl.value == s4
        return out

    def transcribe_streaming(self, mel, t_embed=None, timings: Timings | None = None):
        """model.rs:873-963.  mel [B,128,T] (or [128,T]); returns list of token ids for B==1
        input given as 2-D/3-D with B==1, else an int32 array [B, n]."""
        mel = self._mel3(mel)
        b, _, t = mel.shape
        cap = b * max(t // 16 + 2, 1)
        out = np.zeros(cap, np.int32)
        n = C.c_int32()
        tm = _Timings()
        _check(lib().vox_transcribe_streaming(self._s, _ptr(mel), b, t, _ptr(out), out.size, C.byref(n), C.byref(tm)))
        self._fill(timings, tm)
        ids = out[: b * n.value].reshape(b, n.value)
        return ids[0].tolist() if b == 1 else ids

    def transcribe_pcm(self, samples, peak_normalize: bool = True, timings: Timings | None = None) -> np.ndarray:
        """Full pipeline for B equal-length streams: samples [B,n] (or [n]) host f32 -> ids [B, n_out]."""
        s = _f32(samples)
        if s.ndim == 1:
            s = s[None]
        b, n = s.shape
        cap = b * (n // 1280 + 120)
        out = np.zeros(cap, np.int32)
        no = C.c_int32()
        tm = _Timings()
        _check(lib().vox_transcribe_pcm(self._s, _ptr(s), b, n, 1 if peak_normalize else 0, _ptr(out), out.size,
                                        C.byref(no), C.byref(tm)))
        self._fill(timings, tm)
        return out[: b * no.value].reshape(b, no.value)

    def transcribe_pcm_dev(self, samples_dev: DeviceBuffer, b: int, n: int, timings: Timings | None = None) -> np.ndarray:
        cap = b * (n // 1280 + 120)
        out = np.zeros(cap, np.int32)
        no = C.c_int32()
        tm = _Timings()
        _check(lib().vox_transcribe_pcm_dev(self._s, samples_dev.ptr, b, n, _ptr(out), out.size, C.byref(no),
                                            C.byref(tm)))
        self._fill(timings, tm)
        return out[: b * no.value].reshape(b, no.value)

    @staticmethod
    def _fill(timings, tm):
        if timings is not None:
            for f, _ in _Timings._fields_:
                setattr(timings, f, getattr(tm, f))

    def generate_step_with_cache(self, token_ids) -> np.ndarray:
        """ids [B,M] -> logits [B,M,vocab]; appends to the decoder KV cache (model.rs:857-867)."""
        ids = np.ascontiguousarray(token_ids, dtype=np.int32)
        if ids.ndim == 1:
            ids = ids[None]
        b, m = ids.shape
        out = np.empty((b, m, self.info["vocab"]), np.float32)
        _check(lib().vox_generate_step_with_cache(self._s, _ptr(ids), b, m, _ptr(out), out.size))
        return out

    def forward_streaming(self, mel, token_ids) -> np.ndarray:
        """mel [B,128,T] (or [128,T]), ids [B,S] -> logits [B,S,vocab]: teacher-forced full pass with
        inputs = audio_embeds + embed(ids) (model.rs:801-814)."""
        mel = np.ascontiguousarray(mel, np.float32)
        if mel.ndim == 2:
            mel = mel[None]
        ids = np.ascontiguousarray(token_ids, dtype=np.int32)
        if ids.ndim == 1:
            ids = ids[None]
        b, _, t = mel.shape
        assert ids.shape[0] == b
        out = np.empty((b, ids.shape[1], self.info["vocab"]), np.float32)
        _check(lib().vox_forward_streaming(self._s, _ptr(mel), b, t, _ptr(ids), ids.shape[1], _ptr(out), out.size))
        return out

    def prefill(self, token_ids, add_audio: bool = True) -> np.ndarray:
        """ids [B,M] -> next token per stream [B]; greedy argmax on the device (vox_prefill)."""
        ids = np.ascontiguousarray(token_ids, dtype=np.int32)
        if ids.ndim == 1:
            ids = ids[None]
        b, m = ids.shape
        nxt = np.empty(b, np.int32)
        _check(lib().vox_prefill(self._s, _ptr(ids), b, m, int(add_audio), _ptr(nxt)))
        return nxt

    def decode_step(self, tok=None, batch: int | None = None, add_audio: bool = True, read: bool = True):
        """One decode position.  tok None = device-side feedback of the previous call (vox_decode_step)."""
        if tok is not None:
            tok = np.ascontiguousarray(tok, dtype=np.int32).reshape(-1)
            batch = tok.size
        assert batch is not None
        nxt = np.empty(batch, np.int32) if read else None
        _check(lib().vox_decode_step(self._s, _ptr(tok) if tok is not None else None, batch, int(add_audio),
                                     _ptr(nxt) if read else None))
        return nxt

    def cache_len(self) -> int:
        v = C.c_int32()
        _check(lib().vox_session_cache_len(self._s, C.byref(v)))
        return v.value

    def reset_cache(self):
        _check(lib().vox_session_reset(self._s))

    def debug(self, what: str) -> np.ndarray | None:
        n = C.c_size_t()
        _check(lib().vox_session_debug_read(self._s, what.encode(), None, 0, C.byref(n)))
        if n.value == 0:
            return None
        out = np.empty(n.value, np.float32)
        _check(lib().vox_session_debug_read(self._s, what.encode(), _ptr(out), out.size, C.byref(n)))
        return out

    def launch_count(self) -> int:
        v = C.c_uint64()
        _check(lib().vox_session_launch_count(self._s, C.byref(v)))
        return v.value

    def close(self):
        if getattr(self, "_s", None):
            lib().vox_session_free(self._s)
            self._s = None
        if getattr(self, "_m", None):
            lib().vox_model_free(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _StreamStats(C.Structure):
    _fields_ = [("gpu_ms", C.c_float), ("live_sessions", C.c_int32), ("mel_frames", C.c_int32), ("encoder_rows", C.c_int32),
                ("prefills", C.c_int32), ("decode_steps", C.c_int32), ("decode_rows", C.c_int32)]


class StreamingPool:
    """Live streaming sessions on one GPU worker (vox_stream_*).  `model` stays owned by the caller
    and must outlive the pool."""

    def __init__(self, model: "Q4VoxtralModel", max_sessions: int = 8, max_seconds: float = 30.0):
        self._model = model
        self._p = _P()
        _check(lib().vox_stream_pool_create(model._m, max_sessions, max_seconds, C.byref(self._p)))
        self.dec_dim = model.info["dec_dim"]

    def open(self) -> int:
        v = C.c_int32()
        _check(lib().vox_stream_open(self._p, C.byref(v)))
        return v.value

    def push(self, session: int, samples):
        s = _f32(samples).reshape(-1)
        _check(lib().vox_stream_push_pcm(self._p, session, _ptr(s), s.size))

    def finish(self, session: int):
        _check(lib().vox_stream_finish(self._p, session))

    def tick(self) -> dict:
        st = _StreamStats()
        _check(lib().vox_stream_tick(self._p, C.byref(st)))
        return {f[0]: getattr(st, f[0]) for f in _StreamStats._fields_}

    def poll(self, session: int, cap: int = 4096):
        ids = np.empty(cap, np.int32)
        n, done = C.c_size_t(), C.c_int32()
        _check(lib().vox_stream_poll_ids(self._p, session, _ptr(ids), cap, C.byref(n), C.byref(done)))
        return ids[:n.value].tolist(), bool(done.value)

    def audio_embeds(self, session: int) -> np.ndarray:
        n = C.c_int32()
        _check(lib().vox_stream_audio_embeds(self._p, session, None, 0, C.byref(n)))
        out = np.empty((n.value, self.dec_dim), np.float32)
        _check(lib().vox_stream_audio_embeds(self._p, session, _ptr(out), out.size, C.byref(n)))
        return out

    def encode_audio_with_cache(self, session: int, mel) -> np.ndarray:
        """mel chunk [128,T] (or [1,128,T]) -> the chunk's audio embeddings [T/16, dec_dim]; the session's encoder
        K/V caches are extended (Q4VoxtralModel::encode_audio_with_cache, model.rs:790-799)."""
        mel = _f32(mel)
        if mel.ndim == 3:
            mel = mel[0]
        t = mel.shape[1]
        t1 = (t + 2 - 3) // 2 + 1
        s = (t1 + 2 - 3) // 2 + 1
        out = np.empty((s // 4 + 1, self.dec_dim), np.float32)
        n = C.c_int32()
        _check(lib().vox_stream_encode_chunk(self._p, session, _ptr(mel), t, _ptr(out), out.size, C.byref(n)))
        return out[:n.value].copy()

    def close_session(self, session: int):
        _check(lib().vox_stream_close(self._p, session))

    def close(self):
        if getattr(self, "_p", None):
            lib().vox_stream_pool_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Q4ModelLoader:
    """Q4ModelLoader (loader.rs:76-128): from_file / from_bytes / from_shards, then load(device)."""

    def __init__(self, reader: GgufReader | None = None, path: str | None = None):
        self._reader, self._path = reader, path

    @staticmethod
    def from_file(path: str) -> "Q4ModelLoader":
        if not os.path.exists(path):
            raise VoxtralError(2, f"Failed to open {path}")
        return Q4ModelLoader(GgufReader.open(path), path)

    @staticmethod
    def from_bytes(data) -> "Q4ModelLoader":
        return Q4ModelLoader(GgufReader.from_bytes(data))

    @staticmethod
    def from_shards(shards) -> "Q4ModelLoader":
        return Q4ModelLoader(GgufReader.from_shards(shards))

    def load(self, device: int = 0, max_batch: int = 1, max_mel_frames: int = 3000) -> Q4VoxtralModel:
        h = _P()
        _check(lib().vox_model_load_gguf_handle(self._reader._h, device, C.byref(h)))
        return Q4VoxtralModel(h, device, max_batch, max_mel_frames)


# ------------------------------------------------------------------------------ tokenizer
class VoxtralTokenizer:
    """VoxtralTokenizer (tokenizer/mod.rs:56-214), decode-only."""

    def __init__(self, handle):
        self._h = handle

    @staticmethod
    def from_file(path: str) -> "VoxtralTokenizer":
        h = _P()
        _check(lib().vox_tokenizer_from_file(os.fsencode(path), C.byref(h)))
        return VoxtralTokenizer(h)

    @staticmethod
    def from_json(json_str: str) -> "VoxtralTokenizer":
        raw = json_str.encode("utf-8")
        h = _P()
        _check(lib().vox_tokenizer_from_json(raw, len(raw), C.byref(h)))
        return VoxtralTokenizer(h)

    def decode(self, ids) -> str:
        a = np.ascontiguousarray(ids, dtype=np.uint32)
        n = C.c_size_t()
        _check(lib().vox_tokenizer_decode(self._h, _ptr(a), a.size, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value + 1)
        _check(lib().vox_tokenizer_decode(self._h, _ptr(a), a.size, buf, n.value + 1, C.byref(n)))
        return buf.raw[: n.value].decode("utf-8")

    def decode_token(self, token_id: int):
        n, found = C.c_size_t(), C.c_int32()
        _check(lib().vox_tokenizer_decode_token(self._h, token_id, None, 0, C.byref(n), C.byref(found)))
        if not found.value:
            return None
        buf = C.create_string_buffer(n.value + 1)
        _check(lib().vox_tokenizer_decode_token(self._h, token_id, buf, n.value + 1, C.byref(n), C.byref(found)))
        return buf.raw[: n.value].decode("utf-8")

    def vocab_size(self) -> int:
        n = C.c_size_t()
        _check(lib().vox_tokenizer_vocab_size(self._h, C.byref(n)))
        return n.value

    def __del__(self):
        try:
            if self._h:
                lib().vox_tokenizer_free(self._h)
                self._h = None
        except Exception:
            pass
