"""Real-weights bring-up: skipped unless the real artefacts are present -- no checkpoint, tokenizer or audio
exists offline.  Drop the files in place (or point the VOX_REAL_* variables at them) and these become the
end-to-end goldens the reference documents:
  * docs/VOXTRAL_ARCHITECTURE.md:512-516: test_data/mary_had_lamb.wav (15.95 s) ->
    " I spoke in the original phonograph. A little piece of practical poetry"
  * src/tokenizer/mod.rs:255-268: ids [1362, 19135, 1294, 1278, 4618, 40307, 3910, 1046] -> " I spoke in the original phonograph."
"""
import os

import numpy as np
import pytest

GGUF = os.environ.get("VOX_REAL_GGUF", "models/voxtral-q4.gguf")
TOK = os.environ.get("VOX_REAL_TOKENIZER", "models/voxtral/tekken.json")
WAV = os.environ.get("VOX_REAL_WAV", "test_data/mary_had_lamb.wav")
need = [p for p in (GGUF, TOK, WAV) if not os.path.exists(p)]
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(bool(need), reason=f"real artefacts absent: {need}")]


def _wav():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    from eval_wer import read_wav
    return read_wav(WAV)


def test_real_model_golden_transcript(vx):
    tok = vx.VoxtralTokenizer.from_file(TOK)
    assert tok.decode([1362, 19135, 1294, 1278, 4618, 40307, 3910, 1046]) == " I spoke in the original phonograph."
    model = vx.Q4ModelLoader.from_file(GGUF).load(0, max_batch=1, max_mel_frames=3000)
    try:
        audio = vx.peak_normalize(_wav())
        ids = model.transcribe_pcm(audio, peak_normalize=False)[0]
        text = tok.decode([int(t) for t in ids if t >= 1000])
        assert text.startswith(" I spoke in the original phonograph. A little piece of practical poetry"), text
        # the streaming session must give the same ids
        pool = vx.StreamingPool(model, max_sessions=1, max_seconds=20.0)
        sid = pool.open()
        got = []
        for p in range(0, audio.size, 1280):
            pool.push(sid, audio[p:p + 1280]); pool.tick(); got += pool.poll(sid)[0]
        pool.finish(sid); pool.tick(); got += pool.poll(sid)[0]
        assert got == np.asarray(ids).tolist()
        pool.close()
    finally:
        model.close()
