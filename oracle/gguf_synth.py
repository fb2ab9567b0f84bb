"""GGUF *reader* for the oracle (test infrastructure) + re-export of the synthetic-weight writer.

The reader restates GgufReader (reference src/gguf/reader.rs:88-223): magic 0x46554747, version 2|3,
u64 tensor_count, u64 kv_count, KVs (skip_gguf_value 327-376), tensor index (name, ndims, u64 dims[],
u32 dtype, u64 offset), data section at the next 32-byte boundary with offsets relative to it
(177-179, 217-219); dims reversed to PyTorch order (loader.rs:497-499); F32/F16 -> f32
(loader.rs:443-474).  The writer / synthetic generators live in
`voxtral_mini_realtime_rs_b200.synth` (pure data generation shared with bench.py) and are
re-exported here for the tests.
"""
from __future__ import annotations

import mmap
import struct

import numpy as np

from voxtral_mini_realtime_rs_b200.synth import (  # noqa: F401  (re-exports)
    ADAPTER, ALIGN, ENC, F16_T, F32_T, FINAL_NORM, GGUF_MAGIC, Q4_0_T, TOK_EMB, VoxtralConfig,
    build_gguf_bytes, header_bytes, nbytes_of, random_q4_blocks, synth_tensor_bytes, tensor_manifest,
    write_synthetic_gguf,
)

_nbytes = nbytes_of


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
            fm = {0: "B", 1: "b", 2: "H", 3: "h", 4: "I", 5: "i", 6: "f", 7: "B", 10: "Q", 11: "q", 12: "d"}
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
        nb = nbytes_of(dt, shape)
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
        """Dims: `voxtral.*` KVs when present, else the reference defaults (config.rs:441-486);
        remaining dims inferred from tensor shapes."""
        c = VoxtralConfig()
        g = lambda k, dflt: int(self.kv.get(k, dflt))  # noqa: E731
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
