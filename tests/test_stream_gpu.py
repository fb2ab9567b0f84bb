"""CUDA streaming sessions (vox_stream_*, csrc/stream.cu) against the incremental oracle (oracle/streaming.py ==
offline transcribe_streaming, tests/test_oracle_streaming.py): audio pushed in small and ragged pieces, sessions
opened at different times sharing one pool (continuous batching: one encoder row batch and one decoder step per tick
for all of them, paged decoder KV), tokens available before the end of the audio, and -- on the tiny model, whose
sliding window (20) is far smaller than the stream -- the encoder K/V ring overwriting keys older than the window.
"""
import numpy as np
import pytest

from oracle import mel as omel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tiny_model(vx, tiny_gguf):
    m = vx.Q4ModelLoader.from_file(tiny_gguf).load(0, max_batch=1, max_mel_frames=1500)
    yield m
    m.close()


def _offline(oracle, audio):
    info = {}
    ids = oracle.transcribe_streaming(omel.mel_tensor_from_audio(audio), omel.time_embedding(6.0, oracle.cfg.dec_dim), info=info)
    return ids, info


def test_single_session_equals_offline_and_streams(vx, tiny_model, tiny_oracle):
    audio = omel.peak_normalize(omel.speechlike(4.0, 21))
    want, info = _offline(tiny_oracle, audio)
    assert tiny_model.transcribe_pcm(audio, peak_normalize=False)[0].tolist() == want
    for piece in (1280, 3001, 16000):
        pool = vx.StreamingPool(tiny_model, max_sessions=2, max_seconds=8.0)
        sid = pool.open()
        got, before_end = [], 0
        for a in range(0, audio.size, piece):
            pool.push(sid, audio[a:a + piece])
            st = pool.tick()
            ids, done = pool.poll(sid)
            got += ids
            assert not done
        before_end = len(got)
        pool.finish(sid)
        pool.tick()
        ids, done = pool.poll(sid)
        got += ids
        assert done
        assert got == want, (piece, got, want, min(info["margins"]))
        emb = pool.audio_embeds(sid)
        ref = info["audio_embeds"].numpy()
        assert emb.shape == ref.shape and np.abs(emb - ref).max() < 1e-3
        assert 0 < before_end < len(want) and len(want) - before_end <= 12     # tokens really stream
        pool.close_session(sid)
        pool.close()


@pytest.mark.parametrize("n_sessions", [3, 8])
def test_sessions_of_different_ages_share_the_pool(vx, tiny_model, tiny_oracle, n_sessions):
    """Sessions start 0.4 s apart, have different lengths and are fed 80 ms per tick: every tick batches the encoder
    rows and the decoder step of whoever is ready (rows at different positions, KV pages from one pool)."""
    rng = np.random.default_rng(5)
    audios = [omel.peak_normalize(omel.speechlike(float(rng.uniform(2.0, 4.5)), 100 + i)) for i in range(n_sessions)]
    wants = [_offline(tiny_oracle, a)[0] for a in audios]
    pool = vx.StreamingPool(tiny_model, max_sessions=n_sessions, max_seconds=8.0)
    start = [4 * i for i in range(n_sessions)]              # tick at which each session opens (a session steps every 2nd tick)
    sids = [None] * n_sessions
    fed = [0] * n_sessions
    got = [[] for _ in range(n_sessions)]
    finished = [False] * n_sessions
    max_rows = 0
    for tick in range(400):
        for i in range(n_sessions):
            if tick == start[i]:
                sids[i] = pool.open()
            if sids[i] is None or finished[i]:
                continue
            if fed[i] < audios[i].size:
                pool.push(sids[i], audios[i][fed[i]:fed[i] + 1280])
                fed[i] += 1280
            elif fed[i] >= audios[i].size:
                pool.finish(sids[i])
                finished[i] = True
        st = pool.tick()
        max_rows = max(max_rows, st["decode_rows"] // max(1, st["decode_steps"]))
        all_done = True
        for i in range(n_sessions):
            if sids[i] is None:
                all_done = False
                continue
            ids, done = pool.poll(sids[i])
            got[i] += ids
            all_done = all_done and done
        if all_done:
            break
    for i in range(n_sessions):
        assert got[i] == wants[i], (i, got[i], wants[i])
    assert max_rows >= min(n_sessions, 3)      # decoder steps really were shared
    pool.close()


def test_pool_capacity_and_errors(vx, tiny_model):
    pool = vx.StreamingPool(tiny_model, max_sessions=1, max_seconds=2.0)
    sid = pool.open()
    with pytest.raises(vx.VoxtralError, match="in use"):
        pool.open()
    with pytest.raises(vx.VoxtralError, match="max_seconds"):
        pool.push(sid, np.zeros(16000 * 3, np.float32))
    pool.finish(sid)
    with pytest.raises(vx.VoxtralError, match="already finished"):
        pool.push(sid, np.zeros(10, np.float32))
    pool.tick()
    ids, done = pool.poll(sid)
    # no audio at all: the stream is the 76 + 17 padding tokens = 119 040 samples -> 46 positions -> 8 ids (model.rs:883-963)
    assert done and len(ids) == vx.stream_progress(119040, True)[4] == 8
    pool.close_session(sid)
    sid2 = pool.open()
    assert sid2 == sid
    pool.close()


def test_encode_audio_with_cache_matches_upstream_semantics(vx, tiny_model, tiny_oracle):
    """vox_stream_encode_chunk == the oracle's encode_audio_with_cache (model.rs:790-799: chunk-local conv stem, encoder
    K/V caches with RoPE / mask offsets): two chunks, the second one's queries reaching back over a window (20) that is
    smaller than the cached length; and a single chunk == the uncached encode_audio."""
    import torch
    mel = omel.mel_tensor_from_audio(omel.peak_normalize(omel.speechlike(4.0, 6)))
    pool = vx.StreamingPool(tiny_model, max_sessions=2, max_seconds=8.0)
    sid = pool.open()
    cache = tiny_oracle.new_encoder_cache()
    for a, b in ((0, 160), (160, 480), (480, mel.shape[2])):
        exp = tiny_oracle.encode_audio_with_cache(mel[:, :, a:b], cache).numpy()
        got = pool.encode_audio_with_cache(sid, mel[0, :, a:b])
        assert got.shape == exp.shape and np.abs(got - exp).max() < 1e-3, (a, b, np.abs(got - exp).max())
    sid2 = pool.open()
    whole = pool.encode_audio_with_cache(sid2, mel[0])
    assert np.abs(whole - tiny_model.encode_audio(mel)[0]).max() < 1e-4
    with pytest.raises(vx.VoxtralError, match="do not mix"):
        pool.push(sid2, np.zeros(1280, np.float32)); pool.tick(); pool.encode_audio_with_cache(sid2, mel[0, :, :64])
    pool.close()
