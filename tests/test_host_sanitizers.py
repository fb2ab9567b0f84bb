"""Host-side C++ under AddressSanitizer + UBSan (GGUF reader, tokenizer, audio plumbing): the code that parses
untrusted files behind the C ABI must reject or survive truncated / corrupted inputs without memory errors.
Mirrors the spirit of the reference's reader / tokenizer error tests (src/gguf/tests.rs:281-325,
src/tokenizer/mod.rs:216-270) at the byte level."""
import os
import shutil
import subprocess

import pytest

from oracle import gguf_synth, tokenizer as otok

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "voxtral_mini_realtime_rs_b200", "csrc")


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_host_code_under_asan_ubsan(tmp_path):
    exe = tmp_path / "drv"
    cmd = ["g++", "-std=c++17", "-g", "-O1", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-fno-omit-frame-pointer", "-I", CSRC, "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "native", "host_sanitizer_driver.cpp"),
           os.path.join(CSRC, "gguf.cpp"), os.path.join(CSRC, "tokenizer.cpp"), os.path.join(CSRC, "audio_host.cpp"),
           "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and ("asan" in r.stderr.lower() or "sanitize" in r.stderr.lower()):
        pytest.skip("toolchain without sanitizer runtimes")
    assert r.returncode == 0, r.stderr[-2000:]
    gguf = tmp_path / "tiny.gguf"
    gguf_synth.write_synthetic_gguf(str(gguf), gguf_synth.VoxtralConfig.tiny(), seed=3)
    tok = tmp_path / "tekken.json"
    tok.write_text(otok.synthetic_tekken_json())
    r = subprocess.run([str(exe), str(gguf), str(tok)], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0"))
    assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-3000:])
    assert "done" in r.stdout and "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr
