"""Oracle-internal consistency of the incremental encoder API (reference model.rs:437-452, 790-799;
Q4Attention::forward_with_cache 125-174; masks with offset masking.rs:50-107).  This is the checker for
the streaming-session row of SURVEY 8(f); the C ABI does not expose it yet, so there is no GPU side."""
import numpy as np
import torch

from oracle import mel as omel


def _mel(seconds, seed):
    return omel.mel_tensor_from_audio(omel.peak_normalize(omel.speechlike(seconds, seed)))


def test_cached_encoder_single_chunk_equals_batch(tiny_oracle):
    mel = _mel(3.0, 5)
    full = tiny_oracle.encode_audio(mel)
    cached = tiny_oracle.encode_audio_with_cache(mel, tiny_oracle.new_encoder_cache())
    assert full.shape == cached.shape
    assert torch.equal(full, cached)  # same operations in the same order


def test_cached_encoder_chunks_extend_the_cache(tiny_oracle):
    """Two chunks through the cache == the uncached layers over the concatenated per-chunk conv
    outputs: the offset RoPE / causal / sliding-window bookkeeping is position-exact.  The tiny model's
    window (20) is smaller than the sequence, so the offset window mask bites."""
    o = tiny_oracle
    mel = _mel(4.0, 6)
    t_split = 160  # mel frames; a multiple of 4 so both chunks give whole encoder frames
    c1, c2 = mel[:, :, :t_split], mel[:, :, t_split:]
    cache = o.new_encoder_cache()
    y1 = o.encoder_forward_with_cache(c1, cache)
    y2 = o.encoder_forward_with_cache(c2, cache)
    assert cache[0]["k"].shape[0] == y1.shape[0] + y2.shape[0] > o.cfg.enc_window
    # reference computation: per-chunk conv stems (as upstream: no carried conv state), then the
    # whole-sequence layers at offset 0
    x = torch.cat([o.conv_downsample(torch.from_numpy(np.ascontiguousarray(c, np.float32)))[0].transpose(0, 1)
                   for c in (c1, c2)], 0).contiguous()
    for i in range(o.cfg.enc_layers):
        x = o.encoder_layer(x, i)
    from oracle.model import ENC, rms_norm
    ref = rms_norm(x, o.f32(f"{ENC}.transformer.norm.weight"), o.cfg.norm_eps)
    got = torch.cat([y1, y2], 0)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
    # and chunking is NOT the same as the whole-utterance stem (zero padding at the cut): the streaming
    # session of SURVEY 8(f)-1 has to carry conv state to reproduce transcribe_streaming exactly
    whole = o.encoder_forward(mel)
    assert whole.shape == ref.shape
    assert (whole - ref).abs().max().item() > 1e-4


def test_streaming_session_equals_offline_transcribe(tiny_oracle):
    """Feeding the padded utterance 80 ms at a time (and in ragged pieces) through the incremental pipeline of
    oracle/streaming.py yields the ids of the whole-utterance transcribe_streaming (model.rs:873-963), every
    token as soon as its inputs are final: the carried state / lookahead table in that module is sufficient."""
    from oracle.streaming import StreamingOracle
    o = tiny_oracle
    audio = omel.peak_normalize(omel.speechlike(4.0, 21))
    t_embed = omel.time_embedding(6.0, o.cfg.dec_dim)
    info = {}
    offline = o.transcribe_streaming(omel.mel_tensor_from_audio(audio), t_embed, info=info)
    for piece in (1280, 3001):
        st = StreamingOracle(o, t_embed)
        got, emitted_before_end = [], 0
        for a in range(0, audio.size, piece):
            got += st.feed(audio[a:a + piece])
        emitted_before_end = len(got)
        got += st.finish()
        assert got == offline, (piece, got, offline, min(info["margins"]))
        emb = torch.stack(st.audio_embeds)
        assert emb.shape == info["audio_embeds"].shape
        assert (emb - info["audio_embeds"]).abs().max().item() < 2e-5 * max(1.0, info["audio_embeds"].abs().max().item())
        # tokens really stream: most are out before the end of the audio (the rest wait for the right padding)
        assert 0 < emitted_before_end < len(offline)
        assert len(offline) - emitted_before_end <= 12


def test_stream_progress_matches_streaming_oracle(vx, tiny_oracle):
    """vox_stream_progress (host bookkeeping of a streaming session) against the counters of the incremental
    oracle at every 80 ms step and at the end of the stream."""
    from oracle.streaming import StreamingOracle
    o = tiny_oracle
    audio = omel.peak_normalize(omel.speechlike(2.5, 9))
    st = StreamingOracle(o, omel.time_embedding(6.0, o.cfg.dec_dim))

    def check(ended):
        got = vx.stream_progress(st.samples.size, ended, o.cfg.reshape_factor, 38)
        assert got == (len(st.mel), len(st.c1), st.c2_count, len(st.audio_embeds), len(st.ids)), (st.samples.size, ended, got)

    check(False)
    for a in range(0, audio.size, 1280):
        st.feed(audio[a:a + 1280])
        check(False)
    st.finish()
    check(True)
    # whole-utterance counts of the 16 s configuration (375 040 samples -> 2344 / 1172 / 586 / 146 / 108)
    assert vx.stream_progress(375040, True) == (2344, 1172, 586, 146, 108)
    # audio embedding p needs samples up to 2560 p + 2600
    for p in (0, 1, 37, 100):
        need = 2560 * p + 2600
        assert vx.stream_progress(need, False)[3] == p + 1 and vx.stream_progress(need - 1, False)[3] == p


def test_encoder_cache_eviction_keeps_results(tiny_oracle):
    """Dropping keys older than the sliding window (absolute positions kept for RoPE and masks) does not change
    the encoder output: the bounded-memory KV ring a long streaming session needs.  Fed frame-by-frame-ish
    (8 mel frames = 2 encoder frames per call) so that eviction happens many times."""
    o = tiny_oracle
    mel = _mel(6.0, 33)                                   # ~ 186 encoder frames >> window 20
    t = (mel.shape[2] // 8) * 8
    keep, evict = o.new_encoder_cache(), o.new_encoder_cache(evict=True)
    outs_k, outs_e = [], []
    for a in range(0, t, 8):
        outs_k.append(o.encoder_forward_with_cache(mel[:, :, a:a + 8], keep))
        outs_e.append(o.encoder_forward_with_cache(mel[:, :, a:a + 8], evict))
    yk, ye = torch.cat(outs_k), torch.cat(outs_e)
    assert yk.shape == ye.shape and yk.shape[0] > 4 * o.cfg.enc_window
    assert (yk - ye).abs().max().item() < 2e-5 * max(1.0, yk.abs().max().item())
    assert keep[0]["k"].shape[0] == yk.shape[0]
    assert evict[0]["k"].shape[0] <= o.cfg.enc_window + 2   # bounded: window + the chunk being processed
    assert evict[0]["base"] + evict[0]["k"].shape[0] == yk.shape[0]
