"""CPU tests of the boundary: the library loads, exports every symbol include/voxtral.h declares,
and its host-side logic (GGUF reader, audio plumbing, time embedding, tokenizer) matches the oracle.
No compute entry point is exercised here (no GPU); they must fail loudly instead of falling back.
"""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import gguf_synth, mel as omel, q4 as oq4, tokenizer as otok

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "voxtral.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vox_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(vx):
    raw = ctypes.CDLL(vx.lib_path())
    names = _declared_symbols()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, missing
    # and the Python mirror binds each of them with a signature
    from voxtral_mini_realtime_rs_b200 import api
    assert sorted(api._SIGS) == names


def test_version_and_device_count(vx):
    assert vx.lib().vox_version() >= 100
    assert vx.device_count() >= 0


def test_no_cpu_fallback(vx, have_gpu, tiny_gguf):
    if have_gpu:
        pytest.skip("GPU present")
    with pytest.raises(vx.VoxtralError, match="no CUDA device"):
        vx.Q4ModelLoader.from_file(tiny_gguf).load(0)
    with pytest.raises(vx.VoxtralError, match="no CUDA device"):
        vx.MelSpectrogram(0)
    raw = oq4.quantize_f32_to_q4_0(np.ones(32 * 32, np.float32))
    with pytest.raises(vx.VoxtralError, match="no CUDA device"):
        vx.Q4Tensor.from_q4_bytes(raw, (32, 32))


def test_product_does_not_import_oracle():
    import subprocess
    import sys
    code = ("import sys; import voxtral_mini_realtime_rs_b200 as v; v.lib(); "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'")
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "voxtral_mini_realtime_rs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f


# ------------------------------------------------------------------ GGUF reader (reader.rs tests)
def test_gguf_reader_file_bytes_shards(vx, tiny_gguf):
    g = gguf_synth.GgufFile(tiny_gguf)
    r = vx.GgufReader.open(tiny_gguf)
    assert r.version() == 3 and r.tensor_count() == g.tensor_count() == 57
    assert sorted(r.tensor_names()) == sorted(g.tensors)
    data = open(tiny_gguf, "rb").read()
    shards = vx.GgufReader.from_shards([data[:777], data[777:100001], data[100001:]])
    whole = vx.GgufReader.from_bytes(data)
    for name in ("norm.weight", "layers.1.attention.wk.weight", gguf_synth.TOK_EMB,
                 f"{gguf_synth.ENC}.conv_layers.1.conv.weight"):
        dt, shape, _ = g.info(name)
        info = r.tensor_info(name)
        assert info["dtype"] == dt and tuple(reversed(info["shape"])) == shape
        for rd in (r, shards, whole):
            assert np.array_equal(rd.tensor_data(name), g.raw(name)), name
    assert r.tensor_info("nonexistent") is None
    with pytest.raises(vx.VoxtralError, match="not found"):
        r.tensor_data("nonexistent")


def test_gguf_reference_style_builders(vx):
    """tests.rs:281-325 run against the C reader."""
    n = 32 * 64
    raw = oq4.quantize_f32_to_q4_0(np.sin(np.arange(n, dtype=np.float32) * np.float32(0.001) - 1).astype(np.float32))
    r = vx.GgufReader.from_bytes(gguf_synth.build_gguf_bytes([("test.weight", 2, (64, 32), raw)]))
    assert r.version() == 3 and r.tensor_count() == 1
    info = r.tensor_info("test.weight")
    assert info["shape"] == (32, 64) and info["dtype"] == 2
    assert np.array_equal(r.tensor_data("test.weight"), raw)
    a = oq4.quantize_f32_to_q4_0(np.full(1024, 0.1, np.float32))
    b = oq4.quantize_f32_to_q4_0(np.full(2048, 0.2, np.float32))
    r3 = vx.GgufReader.from_bytes(gguf_synth.build_gguf_bytes([("weight_a", 2, (32, 32), a), ("weight_b", 2, (32, 64), b)]))
    assert r3.tensor_count() == 2 and r3.tensor_info("weight_b") is not None and r3.tensor_info("zzz") is None
    assert np.array_equal(r3.tensor_data("weight_b"), b)


def test_gguf_v2_and_kv_types(vx):
    import io
    import struct
    b = io.BytesIO()
    b.write(struct.pack("<IIQQ", 0x46554747, 2, 1, 6))

    def s(x):
        e = x.encode()
        b.write(struct.pack("<Q", len(e)) + e)
    s("a.u8"); b.write(struct.pack("<IB", 0, 7))
    s("a.f32"); b.write(struct.pack("<If", 6, 1.5))
    s("a.str"); b.write(struct.pack("<I", 8)); s("hello")
    s("a.arr"); b.write(struct.pack("<IIQ", 9, 5, 3) + struct.pack("<iii", 1, 2, 3))
    s("a.arr_str"); b.write(struct.pack("<IIQ", 9, 8, 2)); s("x"); s("yz")
    s("a.f64"); b.write(struct.pack("<Id", 12, 2.5))
    s("t"); b.write(struct.pack("<IQIQ", 1, 4, 0, 0))
    b.write(b"\0" * ((-b.tell()) % 32))
    b.write(np.arange(4, dtype=np.float32).tobytes())
    r = vx.GgufReader.from_bytes(b.getvalue())
    assert r.version() == 2
    assert np.array_equal(r.tensor_data("t").view(np.float32), np.arange(4, dtype=np.float32))


def test_gguf_errors(vx, tmp_path):
    with pytest.raises(vx.VoxtralError, match="Invalid GGUF magic"):
        vx.GgufReader.from_bytes(b"NOPE" + b"\0" * 64)
    import struct
    with pytest.raises(vx.VoxtralError, match="Unsupported GGUF version"):
        vx.GgufReader.from_bytes(struct.pack("<IIQQ", 0x46554747, 7, 0, 0))
    with pytest.raises(vx.VoxtralError, match="Failed to read"):
        vx.GgufReader.from_bytes(struct.pack("<IIQQ", 0x46554747, 3, 5, 0))      # truncated index
    with pytest.raises(vx.VoxtralError, match="Failed to open"):
        vx.GgufReader.open(str(tmp_path / "nope.gguf"))
    bad_dtype = struct.pack("<IIQQ", 0x46554747, 3, 1, 0) + struct.pack("<Q", 1) + b"t" + struct.pack("<IQIQ", 1, 32, 9, 0)
    with pytest.raises(vx.VoxtralError, match="Unsupported GGML dtype"):
        vx.GgufReader.from_bytes(bad_dtype)


# ------------------------------------------------------------------ audio plumbing
@pytest.mark.parametrize("n", [0, 1, 1279, 1280, 5000, 255168, 256000])
def test_pad_audio_matches_oracle(vx, n):
    rng = np.random.default_rng(n)
    a = rng.standard_normal(n).astype(np.float32)
    got, exp = vx.pad_audio(a), omel.pad_audio(a)
    assert got.size == exp.size and np.array_equal(got, exp)
    assert got.size % 1280 == 0


def test_pad_audio_custom_config(vx):
    a = np.ones(3000, np.float32)
    cfg = vx.PadConfig(n_left_pad_tokens=32, extra_right_pad_tokens=5)
    exp = omel.pad_audio(a, omel.PadConfig(n_left_pad_tokens=32, extra_right_pad_tokens=5))
    assert np.array_equal(vx.pad_audio(a, cfg), exp)
    assert cfg.left_pad_samples() == 32 * 1280


def test_peak_normalize_matches_oracle(vx):
    rng = np.random.default_rng(1)
    a = (rng.standard_normal(4096) * 0.1).astype(np.float32)
    assert np.array_equal(vx.peak_normalize(a), omel.peak_normalize(a))
    z = np.zeros(16, np.float32)
    assert np.array_equal(vx.peak_normalize(z), z)
    assert vx.peak_normalize(np.zeros(0, np.float32)).size == 0


def test_chunk_plan_matches_oracle(vx):
    for n, mf, ov in ((0, 1500, 0), (240000, 1500, 0), (240001, 1500, 0), (500000, 1200, 0), (160000, 400, 100)):
        assert vx.chunk_audio(n, mf, ov) == omel.chunk_plan(n, mf, overlap_frames=ov)
    assert vx.needs_chunking(240001) and not vx.needs_chunking(240000)
    with pytest.raises(vx.VoxtralError):
        vx.chunk_audio(1000, 10, 10)


def test_time_embedding_matches_oracle(vx):
    e = vx.TimeEmbedding(4).embed(1.0).ravel()
    assert np.allclose(e, [np.cos(1.0), np.cos(0.01), np.sin(1.0), np.sin(0.01)], atol=1e-6)
    got = vx.TimeEmbedding(3072).embed(6.0)
    assert got.shape == (1, 1, 3072)
    assert np.abs(got.ravel() - omel.time_embedding(6.0, 3072)).max() < 2e-6
    with pytest.raises(vx.VoxtralError):
        vx.TimeEmbedding(3).embed(1.0)


def test_mel_num_frames(vx):
    for n in (0, 160, 16000, 375040, 256000):
        assert vx.MelSpectrogram.num_frames(n) == omel.num_frames(n)


# ------------------------------------------------------------------ tokenizer
def test_tokenizer_matches_oracle(vx, tmp_path):
    js = otok.synthetic_tekken_json(400, 8)
    t, o = vx.VoxtralTokenizer.from_json(js), otok.VoxtralTokenizer.from_json(js)
    assert t.vocab_size() == o.vocab_size
    rng = np.random.default_rng(0)
    for _ in range(20):
        ids = rng.integers(0, 1500, size=40).tolist()
        assert t.decode(ids) == o.decode(ids)
    assert t.decode([]) == ""
    assert t.decode([1000 + 8 + 0xC3]) == "�"
    assert t.decode([1000 + 8 + 0xE2, 1000 + 8 + 0x82, 1000 + 8 + 0xAC]) == "€"
    for i in (0, 3, 7, 8, 900, 1008, 1008 + 259, 1008 + 399, 1000 + 5, 99999):
        assert t.decode_token(i) == o.decode_token(i), i
    p = tmp_path / "tekken.json"
    p.write_text(js)
    assert vx.VoxtralTokenizer.from_file(str(p)).decode([1008 + 72, 1008 + 105]) == "Hi"
    with pytest.raises(vx.VoxtralError, match="Failed to open tokenizer file"):
        vx.VoxtralTokenizer.from_file(str(tmp_path / "missing.json"))
    with pytest.raises(vx.VoxtralError, match="Failed to parse"):
        vx.VoxtralTokenizer.from_json("{not json")
    with pytest.raises(vx.VoxtralError, match="missing field"):
        vx.VoxtralTokenizer.from_json('{"config": {}, "vocab": []}')


def test_tokenizer_unicode_escapes(vx):
    js = ('{"config":{"pattern":"","num_vocab_tokens":2,"default_vocab_size":3,"default_num_special_tokens":1,'
          '"version":"v7"},"vocab":[{"rank":0,"token_bytes":null,"token_str":"<s>","is_control":true},'
          '{"rank":1,"token_bytes":null,"token_str":"caf\\u00e9 \\ud83d\\ude00"},'
          '{"rank":2,"token_bytes":"!!!notbase64","token_str":"fb"}]}')
    t, o = vx.VoxtralTokenizer.from_json(js), otok.VoxtralTokenizer.from_json(js)
    assert t.decode([1001]) == o.decode([1001]) == "café 😀"
    assert t.decode([1002]) == o.decode([1002]) == "fb"      # invalid base64 -> token_str fallback
    assert t.decode_token(0) == "<s>"


def test_rust_binding_names_are_exported(vx):
    """rust/voxtral_sys.rs (source-only FFI binding for the reference crate) declares only functions the library exports."""
    import re
    src = open(os.path.join(ROOT, "rust", "voxtral_sys.rs")).read()
    names = re.findall(r"pub fn (vox_\w+)\(", src)
    assert len(names) > 40
    lib = vx.lib()
    for n in names:
        assert hasattr(lib, n), n
