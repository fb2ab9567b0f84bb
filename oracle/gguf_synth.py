"""GGUF v3 writer/reader + synthetic Voxtral weight generator (test infrastructure).

File layout follows what the reference parses (src/gguf/reader.rs:105-188): magic
0x46554747, version (2|3), u64 tensor_count, u64 kv_count, KVs, tensor index
(name, ndims, u64 dims[], u32 dtype, u64 offset), data section at the next 32-byte
boundary, offsets relative to it (reader.rs:177-179, 217-219).  Dims are stored
reversed w.r.t. PyTorch order (loader.rs:497-499).  dtype codes F32=0, F16=1, Q4_0=2
(reader.rs:28-35).  Tensor names: src/models/weights.rs:219-230,294-390.
The in-memory builders of src/gguf/tests.rs:90-168 (`build_minimal_gguf`,
`build_multi_tensor_gguf`) are restated by `build_gguf_bytes`.

Model dims: the reference hard-codes them (`config.rs:441-486`) and skips all metadata
KVs.  Our synthetic *tiny* models carry optional `voxtral.*` u32 KVs so that the same
loader can run them; the reference would simply skip those keys.
"""
from __future__ import annotations

import io
import math
import mmap
import os
import struct
import zlib
from dataclasses import dataclass, asdict

import numpy as np

from . import q4

GGUF_MAGIC = 0x46554747
ALIGN = 32
F32_T, F16_T, Q4_0_T = 0, 1, 2

ENC = "mm_streams_embeddings.embedding_module.whisper_encoder"
ADAPTER = "mm_streams_embeddings.embedding_module.audio_language_projection"
TOK_EMB = "mm_streams_embeddings.embedding_module.tok_embeddings.weight"
FINAL_NORM = "norm.weight"


@dataclass
class VoxtralConfig:
    # encoder (config.rs:441-460)
    n_mels: int = 128
    enc_dim: int = 1280
    enc_layers: int = 32
    enc_heads: int = 32
    enc_head_dim: int = 64
    enc_ffn: int = 5120
    enc_window: int = 750
    # decoder (config.rs:462-486)
    dec_dim: int = 3072
    dec_layers: int = 26
    dec_heads: int = 32
    dec_kv_heads: int = 8
    dec_head_dim: int = 128
    dec_ffn: int = 9216
    dec_window: int = 8192
    vocab: int = 131072
    t_cond_dim: int = 32
    reshape_factor: int = 4
    rope_theta: float = 1_000_000.0
    norm_eps: float = 1e-5

    @staticmethod
    def tiny() -> "VoxtralConfig":
        """Small config with the same structure (non-square projections, GQA 2:1,
        sliding window small enough to bite)."""
        return VoxtralConfig(enc_dim=64, enc_layers=2, enc_heads=4, enc_head_dim=32, enc_ffn=128,
                             enc_window=20, dec_dim=96, dec_layers=2, dec_heads=4, dec_kv_heads=2,
                             dec_head_dim=32, dec_ffn=160, dec_window=8192, vocab=512)

    def kv_items(self):
        return [
            ("voxtral.enc.n_layers", self.enc_layers), ("voxtral.enc.n_heads", self.enc_heads),
            ("voxtral.enc.head_dim", self.enc_head_dim), ("voxtral.enc.sliding_window", self.enc_window),
            ("voxtral.dec.n_layers", self.dec_layers), ("voxtral.dec.n_heads", self.dec_heads),
            ("voxtral.dec.n_kv_heads", self.dec_kv_heads), ("voxtral.dec.head_dim", self.dec_head_dim),
            ("voxtral.dec.sliding_window", self.dec_window),
            ("voxtral.reshape_factor", self.reshape_factor),
        ]


def tensor_manifest(cfg: VoxtralConfig):
    """[(name, dtype_code, torch_shape)] in file order."""
    out = []
    d, hd = cfg.enc_dim, cfg.enc_heads * cfg.enc_head_dim
    out += [(f"{ENC}.conv_layers.0.conv.weight", F32_T, (d, cfg.n_mels, 3)),
            (f"{ENC}.conv_layers.0.conv.bias", F32_T, (d,)),
            (f"{ENC}.conv_layers.1.conv.weight", F32_T, (d, d, 3)),
            (f"{ENC}.conv_layers.1.conv.bias", F32_T, (d,))]
    for i in range(cfg.enc_layers):
        p = f"{ENC}.transformer.layers.{i}"
        out += [(f"{p}.attention_norm.weight", F32_T, (d,)),
                (f"{p}.attention.wq.weight", Q4_0_T, (hd, d)), (f"{p}.attention.wq.bias", F32_T, (hd,)),
                (f"{p}.attention.wk.weight", Q4_0_T, (hd, d)),
                (f"{p}.attention.wv.weight", Q4_0_T, (hd, d)), (f"{p}.attention.wv.bias", F32_T, (hd,)),
                (f"{p}.attention.wo.weight", Q4_0_T, (d, hd)), (f"{p}.attention.wo.bias", F32_T, (d,)),
                (f"{p}.ffn_norm.weight", F32_T, (d,)),
                (f"{p}.feed_forward.w1.weight", Q4_0_T, (cfg.enc_ffn, d)),
                (f"{p}.feed_forward.w2.weight", Q4_0_T, (d, cfg.enc_ffn)),
                (f"{p}.feed_forward.w2.bias", F32_T, (d,)),
                (f"{p}.feed_forward.w3.weight", Q4_0_T, (cfg.enc_ffn, d))]
    out += [(f"{ENC}.transformer.norm.weight", F32_T, (d,))]
    D = cfg.dec_dim
    out += [(f"{ADAPTER}.0.weight", Q4_0_T, (D, d * cfg.reshape_factor)),
            (f"{ADAPTER}.2.weight", Q4_0_T, (D, D)),
            (TOK_EMB, Q4_0_T, (cfg.vocab, D))]
    qd, kvd = cfg.dec_heads * cfg.dec_head_dim, cfg.dec_kv_heads * cfg.dec_head_dim
    for j in range(cfg.dec_layers):
        p = f"layers.{j}"
        out += [(f"{p}.ada_rms_norm_t_cond.0.weight", Q4_0_T, (cfg.t_cond_dim, D)),
                (f"{p}.ada_rms_norm_t_cond.2.weight", Q4_0_T, (D, cfg.t_cond_dim)),
                (f"{p}.attention_norm.weight", F32_T, (D,)),
                (f"{p}.attention.wq.weight", Q4_0_T, (qd, D)),
                (f"{p}.attention.wk.weight", Q4_0_T, (kvd, D)),
                (f"{p}.attention.wv.weight", Q4_0_T, (kvd, D)),
                (f"{p}.attention.wo.weight", Q4_0_T, (D, qd)),
                (f"{p}.ffn_norm.weight", F32_T, (D,)),
                (f"{p}.feed_forward.w1.weight", Q4_0_T, (cfg.dec_ffn, D)),
                (f"{p}.feed_forward.w2.weight", Q4_0_T, (D, cfg.dec_ffn)),
                (f"{p}.feed_forward.w3.weight", Q4_0_T, (cfg.dec_ffn, D))]
    out += [(FINAL_NORM, F32_T, (D,))]
    return out


def _nbytes(dtype: int, shape) -> int:
    n = int(np.prod(shape))
    return {F32_T: n * 4, F16_T: n * 2, Q4_0_T: n // 32 * 18}[dtype]


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))


def synth_tensor_bytes(name: str, dtype: int, shape, seed: int, mode: str = "blocks") -> np.ndarray:
    """Deterministic synthetic payload (uint8 array) for one tensor.

    Q4 'blocks' mode: uniform random nibbles (all 16 values incl. 0 => -8*d, which the
    reference's test quantiser never emits -- SURVEY quantiser note) and f16 scales
    d = s/4.64 * U(0.5,1.5) so that the dequantised weights have std ~= s.
    Q4 'gauss' mode: N(0,s^2) f32 quantised by the reference test quantiser.
    """
    rng = _rng(seed, name)
    n = int(np.prod(shape))
    if dtype == Q4_0_T:
        k = int(shape[-1])
        s = 1.0 / math.sqrt(k)          # variance-preserving linear
        if name == f"{ADAPTER}.2.weight":
            s = 2.0 / k                 # audio embeds ~ text embeds so token feedback matters
        if mode == "gauss":
            return q4.quantize_f32_to_q4_0((rng.standard_normal(n) * s).astype(np.float32))
        nb = n // 32
        out = np.empty((nb, 18), np.uint8)
        d = (s / 4.3 * rng.uniform(0.5, 1.5, nb)).astype(np.float16)
        out[:, 0:2] = d.view(np.uint8).reshape(nb, 2)
        # nibbles: 1..15 uniformly (q-8 symmetric in -7..7), with ~1/64 forced to 0 (=-8*d)
        a = rng.integers(0, 256, size=(nb, 16), dtype=np.uint8)
        lo = (a % 15 + 1).astype(np.uint8)
        hi = ((a // 15) % 15 + 1).astype(np.uint8)   # a//15 in 0..17
        z = rng.integers(0, 256, size=(nb, 16), dtype=np.uint8)
        lo[z < 4] = 0
        hi[(z >= 4) & (z < 8)] = 0
        out[:, 2:] = lo | (hi << 4)
        return out.reshape(-1)
    if name.endswith("norm.weight"):
        v = 1.0 + 0.01 * rng.standard_normal(n)
    elif name.endswith(".bias"):
        v = 0.01 * rng.standard_normal(n)
    else:  # conv weights [out, in, 3]
        v = (1.5 / math.sqrt(shape[1] * shape[2])) * rng.standard_normal(n)
    if dtype == F16_T:
        return v.astype(np.float16).view(np.uint8)
    return v.astype(np.float32).view(np.uint8)


def _w_str(b: io.BufferedIOBase, s: str):
    e = s.encode()
    b.write(struct.pack("<Q", len(e)))
    b.write(e)


def _header_bytes(tensors, kvs, version=3) -> bytes:
    """tensors: [(name, dtype, torch_shape, offset)]"""
    b = io.BytesIO()
    b.write(struct.pack("<IIQQ", GGUF_MAGIC, version, len(tensors), len(kvs)))
    for key, val in kvs:
        _w_str(b, key)
        if isinstance(val, str):
            b.write(struct.pack("<I", 8))
            _w_str(b, val)
        elif isinstance(val, float):
            b.write(struct.pack("<If", 6, val))
        else:
            b.write(struct.pack("<II", 4, int(val)))
    for name, dtype, shape, off in tensors:
        _w_str(b, name)
        b.write(struct.pack("<I", len(shape)))
        for dim in reversed(shape):  # GGUF order = reversed torch order
            b.write(struct.pack("<Q", int(dim)))
        b.write(struct.pack("<IQ", dtype, off))
    pad = (-b.tell()) % ALIGN
    b.write(b"\0" * pad)
    return b.getvalue()


def build_gguf_bytes(tensors, kvs=(("general.architecture", "voxtral"),), version=3) -> bytes:
    """In-memory GGUF like tests.rs:90-168.  tensors: [(name, dtype, torch_shape, uint8 data)].
    Offsets are cumulative without inter-tensor padding, as in the reference's builders."""
    metas, off = [], 0
    for name, dtype, shape, data in tensors:
        metas.append((name, dtype, shape, off))
        off += len(data)
    out = io.BytesIO()
    out.write(_header_bytes(metas, list(kvs), version))
    for _, _, _, data in tensors:
        out.write(bytes(data))
    return out.getvalue()


def write_synthetic_gguf(path: str, cfg: VoxtralConfig, seed: int = 42, mode: str = "blocks",
                         f16_norms: bool = False) -> dict:
    """Stream a synthetic Voxtral GGUF to `path` (tensor data 32-byte aligned, like
    llama.cpp-produced files).  Returns {"bytes": total, "q4_bytes": ...}."""
    man = tensor_manifest(cfg)
    if f16_norms:
        man = [(n, F16_T if (dt == F32_T and n.endswith("norm.weight")) else dt, sh) for n, dt, sh in man]
    metas, off, q4b = [], 0, 0
    for name, dtype, shape in man:
        off = (off + ALIGN - 1) // ALIGN * ALIGN
        metas.append((name, dtype, shape, off))
        nb = _nbytes(dtype, shape)
        off += nb
        if dtype == Q4_0_T:
            q4b += nb
    kvs = [("general.architecture", "voxtral")] + cfg.kv_items()
    hdr = _header_bytes(metas, kvs)
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(hdr)
        base = f.tell()
        for name, dtype, shape, o in metas:
            cur = f.tell() - base
            if cur < o:
                f.write(b"\0" * (o - cur))
            f.write(synth_tensor_bytes(name, dtype, shape, seed, mode).tobytes())
        total = f.tell()
    os.replace(tmp, path)
    return {"bytes": total, "q4_bytes": q4b, "tensors": len(metas)}


# ---------------------------------------------------------------------------
# Reader (restates GgufReader, reader.rs:88-223) over an mmap or bytes object
# ---------------------------------------------------------------------------
class GgufFile:
    def __init__(self, src):
        if isinstance(src, (bytes, bytearray, memoryview)):
            self.buf = memoryview(src)
            self._f = None
        else:
            self._f = open(src, "rb")
            self.buf = memoryview(mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ))
        self.kv = {}
        self.tensors = {}
        self._parse()

    def _parse(self):
        b, p = self.buf, 0

        def rd(fmt):
            nonlocal p
            v = struct.unpack_from("<" + fmt, b, p)
            p += struct.calcsize("<" + fmt)
            return v if len(v) > 1 else v[0]

        def rstr():
            nonlocal p
            n = rd("Q")
            s = bytes(b[p:p + n]).decode("utf-8")
            p += n
            return s

        def rval(t):
            nonlocal p
            fm = {0: "B", 1: "b", 2: "H", 3: "h", 4: "I", 5: "i", 6: "f", 7: "B", 10: "Q", 11: "q",
                  12: "d"}
            if t in fm:
                return rd(fm[t])
            if t == 8:
                return rstr()
            if t == 9:
                et = rd("I")
                cnt = rd("Q")
                return [rval(et) for _ in range(cnt)]
            raise ValueError(f"Unknown GGUF metadata value type: {t}")

        magic = rd("I")
        if magic != GGUF_MAGIC:
            raise ValueError(f"Invalid GGUF magic: 0x{magic:08X}")
        self.version = rd("I")
        if self.version not in (2, 3):
            raise ValueError(f"Unsupported GGUF version: {self.version}")
        n_t, n_kv = rd("Q"), rd("Q")
        for _ in range(n_kv):
            k = rstr()
            t = rd("I")
            self.kv[k] = rval(t)
        for _ in range(n_t):
            name = rstr()
            nd = rd("I")
            dims = [rd("Q") for _ in range(nd)]
            dt = rd("I")
            if dt not in (0, 1, 2):
                raise ValueError(f"Unsupported GGML dtype code: {dt}")
            off = rd("Q")
            self.tensors[name] = (dt, tuple(reversed(dims)), off)
        self.data_off = (p + ALIGN - 1) // ALIGN * ALIGN

    def tensor_count(self):
        return len(self.tensors)

    def info(self, name):
        return self.tensors.get(name)

    def raw(self, name) -> np.ndarray:
        dt, shape, off = self.tensors[name]
        nb = _nbytes(dt, shape)
        a = self.data_off + off
        return np.frombuffer(self.buf[a:a + nb], dtype=np.uint8)

    def f32(self, name) -> np.ndarray:
        """loader.rs:443-474 (F32 / F16 -> f32)."""
        dt, shape, _ = self.tensors[name]
        raw = self.raw(name)
        if dt == F32_T:
            return raw.view(np.float32).reshape(shape).copy()
        if dt == F16_T:
            return raw.view(np.float16).astype(np.float32).reshape(shape)
        raise ValueError(f"Cannot load Q4_0 tensor '{name}' as f32")

    def config(self) -> VoxtralConfig:
        """Dims: `voxtral.*` KVs when present, else the reference defaults; remaining dims
        inferred from tensor shapes."""
        c = VoxtralConfig()
        kv = self.kv
        g = lambda k, dflt: int(kv.get(k, dflt))
        c.enc_layers = g("voxtral.enc.n_layers", c.enc_layers)
        c.enc_heads = g("voxtral.enc.n_heads", c.enc_heads)
        c.enc_head_dim = g("voxtral.enc.head_dim", c.enc_head_dim)
        c.enc_window = g("voxtral.enc.sliding_window", c.enc_window)
        c.dec_layers = g("voxtral.dec.n_layers", c.dec_layers)
        c.dec_heads = g("voxtral.dec.n_heads", c.dec_heads)
        c.dec_kv_heads = g("voxtral.dec.n_kv_heads", c.dec_kv_heads)
        c.dec_head_dim = g("voxtral.dec.head_dim", c.dec_head_dim)
        c.dec_window = g("voxtral.dec.sliding_window", c.dec_window)
        c.reshape_factor = g("voxtral.reshape_factor", c.reshape_factor)
        c.enc_dim, c.n_mels, _ = self.tensors[f"{ENC}.conv_layers.0.conv.weight"][1]
        c.enc_ffn = self.tensors[f"{ENC}.transformer.layers.0.feed_forward.w1.weight"][1][0]
        c.vocab, c.dec_dim = self.tensors[TOK_EMB][1]
        c.dec_ffn = self.tensors["layers.0.feed_forward.w1.weight"][1][0]
        c.t_cond_dim = self.tensors["layers.0.ada_rms_norm_t_cond.0.weight"][1][0]
        return c
