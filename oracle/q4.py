"""Q4_0 block arithmetic -- oracle (test infrastructure, see oracle/__init__.py).

Restates, in numpy:
  * the reference's *test* quantiser      src/gguf/tests.rs:24-57
    (d = amax/7, q = min(15, trunc_sat_u8(x/d + 8.5)); duplicated in
    tests/gguf_integration.rs:20-44 and benches/q4_ops.rs:16-39)
  * the dequantiser                       src/gguf/tensor.rs:83-113, loader.rs:505-521,
                                          tests.rs:60-87
  * the f32 triple-loop matmul            src/gguf/tests.rs:172-185 (`reference_matmul`)
  * the fused kernel's accumulation order src/gguf/shader.wgsl:96-127 /
                                          shader_naive.wgsl:62-94

Block layout (18 bytes / 32 weights): f16 scale `d` (LE) then 16 bytes; byte i holds
element i in its low nibble and element i+16 in its high nibble; value = (nibble-8)*d.

Pinned by: tests.rs `test_q4_block_dequant` (190-227), `test_q4_block_edge_cases`
(229-274) -- closed-form inputs, reproduced in tests/test_oracle_pins.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import numpy as np

BLOCK = 32
BLOCK_BYTES = 18


def quantize_f32_to_q4_0(data: np.ndarray) -> np.ndarray:
    """Reference test quantiser (tests.rs:24-57).  Returns uint8[(n/32)*18]."""
    x = np.ascontiguousarray(data, dtype=np.float32).reshape(-1)
    assert x.size % BLOCK == 0, f"length {x.size} not a multiple of 32"
    nb = x.size // BLOCK
    blk = x.reshape(nb, BLOCK)
    amax = np.max(np.abs(blk), axis=1).astype(np.float32)
    d = (amax / np.float32(7.0)).astype(np.float32)
    with np.errstate(divide="ignore"):
        inv = np.where(d != 0, np.float32(1.0) / d, np.float32(0.0)).astype(np.float32)
    d16 = d.astype(np.float16)  # half::f16::from_f32 == round-to-nearest-even
    v = (blk * inv[:, None]).astype(np.float32) + np.float32(8.5)
    # Rust `as u8`: truncate toward zero, saturate to [0,255], NaN -> 0
    v = np.nan_to_num(v, nan=0.0)
    q = np.clip(np.trunc(v), 0, 255).astype(np.uint8)
    q = np.minimum(q, 15).astype(np.uint8)
    out = np.empty((nb, BLOCK_BYTES), dtype=np.uint8)
    out[:, 0:2] = d16.view(np.uint8).reshape(nb, 2)
    out[:, 2:] = q[:, :16] | (q[:, 16:] << 4)
    return out.reshape(-1)


def dequantize_q4_0(raw: np.ndarray, n_elements: int | None = None) -> np.ndarray:
    """tensor.rs:83-113.  raw: uint8[(n/32)*18] -> float32[n]."""
    raw = np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1, BLOCK_BYTES)
    nb = raw.shape[0]
    if n_elements is not None:
        assert n_elements == nb * BLOCK
    d = raw[:, 0:2].copy().view(np.float16).astype(np.float32).reshape(nb, 1)
    qs = raw[:, 2:]
    lo = (qs & 0x0F).astype(np.float32) - np.float32(8.0)
    hi = ((qs >> 4) & 0x0F).astype(np.float32) - np.float32(8.0)
    out = np.empty((nb, BLOCK), dtype=np.float32)
    out[:, :16] = lo * d
    out[:, 16:] = hi * d
    return out.reshape(-1)


def reference_matmul(a: np.ndarray, b_t: np.ndarray, m: int, k: int, n: int) -> np.ndarray:
    """tests.rs:172-185: out[i,j] = sum_l a[i,l]*b_t[j,l], sequential f32 accumulation.

    Implemented with a cumulative f32 loop over K (vectorised over i,j) so the
    accumulation order is the reference's, not BLAS's.
    """
    a = np.asarray(a, np.float32).reshape(m, k)
    b = np.asarray(b_t, np.float32).reshape(n, k)
    acc = np.zeros((m, n), np.float32)
    for l in range(k):
        acc += a[:, l : l + 1] * b[None, :, l]
    return acc


def q4_matmul_shader_order(x: np.ndarray, raw: np.ndarray, n: int, k: int) -> np.ndarray:
    """The fused kernel's arithmetic (shader.wgsl:96-127): per block, per word wi,
    acc += dot4((lo-8)*d, x[..]) then acc += dot4((hi-8)*d, x[16+..]).  f32.
    x: [M,K] -> [M,N].  Pure numpy, vectorised over (M,N); small shapes only.
    """
    x = np.asarray(x, np.float32)
    m = x.shape[0]
    w = dequantize_q4_0(raw, n * k).reshape(n, k // BLOCK, BLOCK)
    xb = x.reshape(m, k // BLOCK, BLOCK)
    acc = np.zeros((m, n), np.float32)
    for b in range(k // BLOCK):
        for wi in range(4):
            for half in (0, 16):
                s = slice(half + wi * 4, half + wi * 4 + 4)
                p = xb[:, None, b, s] * w[None, :, b, s]  # [M,N,4]
                dot = ((p[..., 0] + p[..., 1]) + p[..., 2]) + p[..., 3]
                acc += dot.astype(np.float32)
    return acc


# ---------------------------------------------------------------------------
# C restatement (oracle/q4_ref.c) -- used for sizes where numpy loops are too slow
# and as the timed CPU baseline ("port").
# ---------------------------------------------------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c(force: bool = False) -> str:
    """gcc build of the C restatements: q4_ref.c / mel_ref.c strict (-O2, no FMA contraction: the
    checker's arithmetic), q4_fast.c vectorised (-O3 -mavx2 -mfma: the timed CPU baseline)."""
    so = os.path.join(_HERE, "liboracle_ref.so")
    strict = [os.path.join(_HERE, "q4_ref.c"), os.path.join(_HERE, "mel_ref.c")]
    fast = [os.path.join(_HERE, "q4_fast.c")]
    if not force and os.path.exists(so) and all(
        os.path.getmtime(so) >= os.path.getmtime(s) for s in strict + fast
    ):
        return so
    objs = []
    for src in strict:
        o = src[:-2] + ".o"
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-march=native", "-ffp-contract=off", "-c", src, "-o", o])
        objs.append(o)
    for src in fast:
        o = src[:-2] + ".o"
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-fPIC", "-mavx2", "-mfma", "-ffp-contract=fast", "-c", src, "-o", o])
        objs.append(o)
    subprocess.check_call(["gcc", "-shared", "-fopenmp", "-o", so] + objs + ["-lm"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c())
        f = ctypes.c_void_p
        _LIB.oracle_q4_matmul.argtypes = [f, f, f, f, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int]
        _LIB.oracle_q4_matmul.restype = None
        _LIB.oracle_q4_matmul_fast.argtypes = [f, f, f, f, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        _LIB.oracle_q4_matmul_fast.restype = None
        _LIB.oracle_q4_dequant.argtypes = [f, f, ctypes.c_longlong]
        _LIB.oracle_q4_dequant.restype = None
        _LIB.oracle_mel_compute_log.argtypes = [f, ctypes.c_longlong, f]
        _LIB.oracle_mel_compute_log.restype = ctypes.c_longlong
    return _LIB


# bench.py sets this for its timed CPU legs: the vectorised re-association (q4_fast.c) instead of the strict
# shader-order loop (q4_ref.c).  Tests never set it: the checker's arithmetic is the strict one.
FAST = False


def q4_matmul_c(x: np.ndarray, raw: np.ndarray, n: int, k: int, bias: np.ndarray | None = None,
                threads: int = 0) -> np.ndarray:
    """y[M,N] = x[M,K] . Wq4[N,K]^T (+bias) via oracle/q4_ref.c (shader accumulation
    order, f32, OpenMP over N)."""
    x = np.ascontiguousarray(x, np.float32)
    m = x.shape[0]
    assert x.shape[1] == k
    raw = np.ascontiguousarray(raw, np.uint8)
    assert raw.size == n * k // BLOCK * BLOCK_BYTES
    y = np.empty((m, n), np.float32)
    b = None
    if bias is not None:
        b = np.ascontiguousarray(bias, np.float32)
    fn = _lib().oracle_q4_matmul_fast if FAST else _lib().oracle_q4_matmul
    fn(x.ctypes.data, raw.ctypes.data, y.ctypes.data, b.ctypes.data if b is not None else None, m, n, k, threads)
    return y


def dequantize_c(raw: np.ndarray) -> np.ndarray:
    raw = np.ascontiguousarray(raw, np.uint8)
    nb = raw.size // BLOCK_BYTES
    out = np.empty(nb * BLOCK, np.float32)
    _lib().oracle_q4_dequant(raw.ctypes.data, out.ctypes.data, nb)
    return out
