"""Generates the committed golden fixtures with the CPU oracle (run in the build container):

    python tests/golden/make_golden.py [--full]

  tiny_s3.npz      tiny synthetic model (seed 3), 4 s speech-like audio (seed 1234):
                   token ids, top-2 margins, audio embeds, mel of the padded audio.
  full_s42_16s.npz full-size synthetic Voxtral-Mini-4B Q4_0 (seed 42, 'blocks' mode, 2.5 GB GGUF
                   regenerated deterministically on the GPU box -- not committed), 16 s
                   speech-like audio (seed 1234): 108 token ids, margins, selected audio-embed
                   rows + per-row sums.  (--full; takes several minutes of CPU.)

The reference itself cannot run here (no Rust toolchain / weights), so these are
oracle outputs, not reference outputs: they pin the CUDA path to the oracle across machines.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import gguf_synth, mel as omel  # noqa: E402
from oracle.model import OracleModel        # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FULL_ROWS = [0, 1, 37, 38, 39, 100, 144, 145]


def tiny():
    cfg = gguf_synth.VoxtralConfig.tiny()
    path = "/tmp/golden_tiny.gguf"
    gguf_synth.write_synthetic_gguf(path, cfg, seed=3)
    audio = omel.speechlike(4.0, seed=1234)
    mel = omel.mel_tensor_from_audio(omel.peak_normalize(audio))
    om = OracleModel(path)
    info = {}
    toks = om.transcribe_streaming(mel, omel.time_embedding(6.0, cfg.dec_dim), info=info)
    np.savez_compressed(os.path.join(HERE, "tiny_s3.npz"), tokens=np.array(toks, np.int32),
                        margins=np.array(info["margins"], np.float32), second=np.array(info["second"], np.int32),
                        audio_embeds=info["audio_embeds"].numpy().astype(np.float32),
                        mel=mel.astype(np.float32))
    print("tiny:", len(toks), "tokens, min margin", min(info["margins"]))


def full(seconds=16.0, gguf="/dev/shm/voxtral_synth_s42.gguf", audio_seed=1234, out_name="full_s42_16s.npz"):
    cfg = gguf_synth.VoxtralConfig()
    t0 = time.time()
    if not os.path.exists(gguf):
        print(gguf_synth.write_synthetic_gguf(gguf, cfg, seed=42), f"{time.time() - t0:.1f}s")
    audio = omel.speechlike(seconds, seed=audio_seed)
    mel = omel.mel_tensor_from_audio(omel.peak_normalize(audio))
    om = OracleModel(gguf)
    info = {}
    t0 = time.time()
    emb = om.encode_audio(mel)
    print(f"oracle encode {time.time() - t0:.1f}s", emb.shape)
    t0 = time.time()
    toks = om.transcribe_streaming(mel, omel.time_embedding(6.0, cfg.dec_dim), audio_embeds=emb, info=info)
    print(f"oracle decode {time.time() - t0:.1f}s")
    e = emb.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, out_name), audio_seed=np.int32(audio_seed), tokens=np.array(toks, np.int32),
                        margins=np.array(info["margins"], np.float32), second=np.array(info["second"], np.int32), rows=np.array(FULL_ROWS, np.int32),
                        audio_rows=e[FULL_ROWS], row_sums=e.astype(np.float64).sum(1),
                        row_abs_sums=np.abs(e).astype(np.float64).sum(1), seconds=np.float32(seconds))
    print("full:", len(toks), "tokens, distinct", len(set(toks)), "min margin", min(info["margins"]))
    print(toks)


if __name__ == "__main__":
    if "--no-tiny" not in sys.argv:
        tiny()
    if "--full" in sys.argv:
        full()
    if "--full2" in sys.argv:      # second utterance (different audio seed): one golden is thin
        full(audio_seed=99, out_name="full_s42_16s_b.npz")
