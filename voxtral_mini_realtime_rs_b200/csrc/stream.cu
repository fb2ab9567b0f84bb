// stream.cu -- true streaming sessions: audio arrives in arbitrary pieces, every stage advances as far
// as its inputs are final, and every token is emitted as soon as it can be -- with the ids of the whole-utterance
// transcribe_streaming (reference src/gguf/model.rs:873-963).  The reference ships the building blocks but no driver:
//   Q4AudioEncoder::forward_with_cache   model.rs:437-452   (encoder layers over a KV cache)
//   encode_audio_with_cache              model.rs:790-799
//   KVCache::apply_sliding_window        kv_cache.rs:176-203 (unused upstream; it would re-base positions)
// What a session carries (the test-side incremental restatement tests/test_oracle_streaming.py checks derives the table):
//   samples (the padded signal so far) -> log-mel frames (final once samples < 160 i + 200 are known) -> conv1 / conv2
//   frames (k3 s2 p1: output t needs input 2t+1) -> 32 encoder layers over a per-layer K/V RING of window + slack
//   positions (absolute positions for RoPE and the causal / sliding-window masks; keys older than the window are simply
//   overwritten: bounded memory for sessions of any length) -> x4 frame stack + adapter -> one decoder position per
//   160 ms of audio, greedy ids.
// Continuous batching: one pool = one GPU worker.  A tick gathers the new encoder frames of ALL live sessions into one
// row batch (the linears do not care which session a row belongs to; RoPE / ring append / attention take a per-row
// (session, absolute position)), and all sessions that can take a decoder step share ONE decode step -- rows at
// different positions, KV pages from one pool (kernels.h KvView).
#include "stream.h"

#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.h"

namespace vox {

namespace {

constexpr int SA_WARPS = 4, SA_THREADS = SA_WARPS * 32;

// RoPE(q) in place, RoPE(k) and v into the row's session ring at slot (pos % ring).  qkv rows [R][3*HQ].
__global__ void stream_rope_append_kernel(float *__restrict__ qkv, const int ld, const int H, const int hd,
                                          const int *__restrict__ row_slot, const int *__restrict__ row_pos, float *__restrict__ kr,
                                          float *__restrict__ vr, const int ring, const float *__restrict__ cos_t,
                                          const float *__restrict__ sin_t) {
    const int r = blockIdx.x;
    const int slot = row_slot[r], pos = row_pos[r];
    const int half = hd >> 1, HQ = H * hd;
    float *row = qkv + (size_t)r * ld;
    const float *cr = cos_t + (size_t)pos * half, *sr = sin_t + (size_t)pos * half;
    float *kdst = kr + ((size_t)slot * ring + (pos % ring)) * HQ;
    float *vdst = vr + ((size_t)slot * ring + (pos % ring)) * HQ;
    for (int i = threadIdx.x; i < H * half; i += blockDim.x) {
        const int h = i / half, p = i - h * half;
        float *q = row + h * hd + 2 * p;
        const float c = cr[p], s = sr[p];
        const float qr = q[0], qi = q[1];
        q[0] = qr * c - qi * s;
        q[1] = qr * s + qi * c;
        const float *k = row + HQ + h * hd + 2 * p;
        kdst[h * hd + 2 * p] = k[0] * c - k[1] * s;
        kdst[h * hd + 2 * p + 1] = k[0] * s + k[1] * c;
    }
    for (int i = threadIdx.x; i < HQ; i += blockDim.x) vdst[i] = row[2 * HQ + i];
}

// Encoder attention of one (head, row) over the session's K/V ring: keys max(0, pos - window) .. pos (causal + sliding
// window with the cache offset, masking.rs:50-107), online softmax per warp, merged through shared memory.
template <int DPL>
__global__ void __launch_bounds__(SA_THREADS)
stream_enc_attn_kernel(const float *__restrict__ qkv, const int ld, const int H, const int *__restrict__ row_slot,
                       const int *__restrict__ row_pos, const float *__restrict__ kr, const float *__restrict__ vr, const int ring,
                       const int window, const float scale, float *__restrict__ out) {
    constexpr int HD = DPL * 32;
    __shared__ float red_m[SA_WARPS], red_l[SA_WARPS];
    __shared__ float red_acc[SA_WARPS][HD];
    const int h = blockIdx.x, r = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = row_slot[r], pos = row_pos[r];
    const int HQ = H * HD;
    float q[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) q[i] = qkv[(size_t)r * ld + h * HD + lane * DPL + i];
    float m_run = -INFINITY, l_run = 0.0f, acc[DPL];
#pragma unroll
    for (int i = 0; i < DPL; ++i) acc[i] = 0.0f;
    const int j_lo = pos - window > 0 ? pos - window : 0;
    const float *kb = kr + (size_t)slot * ring * HQ + h * HD + lane * DPL;
    const float *vb = vr + (size_t)slot * ring * HQ + h * HD + lane * DPL;
    for (int j = j_lo + warp; j <= pos; j += SA_WARPS) {
        const size_t at = (size_t)(j % ring) * HQ;
        float kk[DPL], vv[DPL];
#pragma unroll
        for (int i = 0; i < DPL; ++i) {
            kk[i] = kb[at + i];
            vv[i] = vb[at + i];
        }
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < DPL; ++i) s = fmaf(q[i], kk[i], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        s *= scale;
        const float m_new = fmaxf(m_run, s);
        const float alpha = expf(m_run - m_new);
        const float p = expf(s - m_new);
        l_run = l_run * alpha + p;
        m_run = m_new;
#pragma unroll
        for (int i = 0; i < DPL; ++i) acc[i] = fmaf(p, vv[i], acc[i] * alpha);
    }
    if (lane == 0) {
        red_m[warp] = m_run;
        red_l[warp] = l_run;
    }
#pragma unroll
    for (int i = 0; i < DPL; ++i) red_acc[warp][lane * DPL + i] = acc[i];
    __syncthreads();
    for (int d = threadIdx.x; d < HD; d += SA_THREADS) {
        float mx = -INFINITY;
#pragma unroll
        for (int w = 0; w < SA_WARPS; ++w) mx = fmaxf(mx, red_m[w]);
        float num = 0.0f, den = 0.0f;
#pragma unroll
        for (int w = 0; w < SA_WARPS; ++w) {
            const float f = (red_m[w] == -INFINITY) ? 0.0f : expf(red_m[w] - mx);
            num = fmaf(red_acc[w][d], f, num);
            den = fmaf(red_l[w], f, den);
        }
        out[(size_t)r * HQ + h * HD + d] = num / den;
    }
}

void launch_stream_attn(const float *qkv, int R, int ld, int H, int hd, const int *row_slot, const int *row_pos, const float *kr,
                        const float *vr, int ring, int window, float scale, float *out, cudaStream_t st) {
    dim3 grid(H, R);
    switch (hd / 32) {
        case 1: stream_enc_attn_kernel<1><<<grid, SA_THREADS, 0, st>>>(qkv, ld, H, row_slot, row_pos, kr, vr, ring, window, scale, out); break;
        case 2: stream_enc_attn_kernel<2><<<grid, SA_THREADS, 0, st>>>(qkv, ld, H, row_slot, row_pos, kr, vr, ring, window, scale, out); break;
        case 4: stream_enc_attn_kernel<4><<<grid, SA_THREADS, 0, st>>>(qkv, ld, H, row_slot, row_pos, kr, vr, ring, window, scale, out); break;
        default: fail(VOX_EINVAL, "stream attention: unsupported head_dim");
    }
    cuda_check(cudaGetLastError(), "stream_enc_attn launch");
}

int conv_out(int t) { return t > 0 ? (t + 2 - 3) / 2 + 1 : 0; }

}  // namespace

// ======================================================================================================
StreamPool *StreamPool::create(Model *m, int max_sessions, float max_seconds) {
    VOX_CHECK(max_sessions >= 1 && max_sessions <= 64, VOX_EINVAL, "max_sessions %d out of range [1,64]", max_sessions);
    VOX_CHECK(max_seconds >= 1.0f && max_seconds <= 60.0f, VOX_EINVAL, "max_seconds %.1f out of range [1,60] (encoder RoPE table: %d frames)",
              max_seconds, m->enc_rope_len);
    const vox_model_info &c = m->info;
    vox_pad_config pc;
    pad_config_default(&pc);
    StreamPool *p = new StreamPool();
    try {
        p->m = m;
        p->pad = pc;
        p->max_sessions = max_sessions;
        p->cap_samples = pad_audio_len((size_t)std::ceil(max_seconds * 16000.0f), pc);
        const int cap_mel = (int)mel_num_frames(p->cap_samples);
        p->s = Session::create(m, max_sessions, cap_mel);
        Session *s = p->s;
        VOX_CHECK(s->S_max <= m->enc_rope_len, VOX_EINVAL, "max_seconds exceeds the encoder RoPE table");
        p->max_new = 256;
        p->ring = c.enc_window + p->max_new;
        const int HQ = c.enc_heads * c.enc_head_dim;
        const size_t B = max_sessions;
        p->pcm = s->arena.alloc_n<float>(B * p->cap_samples);
        p->enc_out = s->arena.alloc_n<float>(B * s->S_max * c.enc_dim);
        const size_t ring_elems = (size_t)c.enc_layers * B * p->ring * HQ;
        p->ek = s->arena.alloc_n<float>(ring_elems);
        p->ev = s->arena.alloc_n<float>(ring_elems);
        const size_t max_rows = (size_t)B * p->max_new;
        p->d_row_slot = s->arena.alloc_n<int>(max_rows);
        p->d_row_pos = s->arena.alloc_n<int>(max_rows);
        p->d_audio_rows = (const float **)s->arena.alloc(sizeof(float *) * B);
        p->slots.resize(max_sessions);
        // decoder KV pages: the session's identity tables are replaced by a free list
        for (int i = s->kv_n_pages - 1; i >= 0; --i) p->free_pages.push_back(i);
        s->stream_mode = true;
    } catch (...) {
        delete p;
        throw;
    }
    return p;
}

StreamPool::~StreamPool() { delete s; }

int StreamPool::open() {
    for (int i = 0; i < max_sessions; ++i)
        if (!slots[i].open) {
            Slot &sl = slots[i];
            sl = Slot();
            sl.open = true;
            // the left padding of pad_audio (pad.rs:89-93) is part of the stream
            sl.n_samples = pad_left(pad);
            CUDA_OK(cudaSetDevice(m->device));
            CUDA_OK(cudaMemsetAsync(pcm + (size_t)i * cap_samples, 0, sizeof(float) * cap_samples, s->st));
            return i;
        }
    fail(VOX_ECAPACITY, fmt("all %d stream sessions are in use", max_sessions));
}

StreamPool::Slot &StreamPool::slot(int id) {
    VOX_CHECK(id >= 0 && id < max_sessions && slots[id].open, VOX_EINVAL, "stream session %d is not open", id);
    return slots[id];
}

void StreamPool::push(int id, const float *samples, size_t n) {
    Slot &sl = slot(id);
    VOX_CHECK(!sl.ended, VOX_EINVAL, "stream session %d already finished", id);
    const size_t worst = sl.n_samples + n + pad_right(pad, sl.n_samples + n);
    VOX_CHECK(worst <= cap_samples, VOX_ECAPACITY, "stream session %d: %zu samples exceed the pool's max_seconds", id, sl.n_audio + n);
    CUDA_OK(cudaSetDevice(m->device));
    if (n) CUDA_OK(cudaMemcpyAsync(pcm + (size_t)id * cap_samples + sl.n_samples, samples, sizeof(float) * n, cudaMemcpyHostToDevice, s->st));
    CUDA_OK(cudaStreamSynchronize(s->st));  // `samples` is caller memory
    sl.n_samples += n;
    sl.n_audio += n;
}

void StreamPool::finish(int id) {
    Slot &sl = slot(id);
    VOX_CHECK(!sl.ended, VOX_EINVAL, "stream session %d already finished", id);
    sl.n_samples += pad_right(pad, sl.n_samples);  // zeros: the buffer was cleared at open()
    sl.ended = true;
}

void StreamPool::close(int id) {
    Slot &sl = slot(id);
    for (int pg : sl.pages) free_pages.push_back(pg);
    sl = Slot();
}

size_t StreamPool::poll(int id, int32_t *ids, size_t cap, bool *done) {
    Slot &sl = slot(id);
    const size_t n = std::min(cap, sl.ids.size() - sl.polled);
    if (n) memcpy(ids, sl.ids.data() + sl.polled, sizeof(int32_t) * n);
    sl.polled += n;
    if (done) *done = sl.ended && sl.drained && sl.polled == sl.ids.size();
    return n;
}

// Encoder layers over `R` gathered rows in s->x_enc (Q4EncoderLayer::forward_with_cache, model.rs:300-315).
void StreamPool::encoder_rows(int R) {
    const vox_model_info &c = m->info;
    const int d = c.enc_dim, HQ = c.enc_heads * c.enc_head_dim;
    const float scale = powf((float)c.enc_head_dim, -0.5f);
    const size_t ring_stride = (size_t)max_sessions * ring * HQ;
    for (int i = 0; i < c.enc_layers; ++i) {
        const EncLayerW &l = m->enc[i];
        s->linear_n(l.wqkv, s->x_enc, R, s->qkv_enc, 3 * HQ, l.bqkv, nullptr, EPI_NONE, l.attn_norm, nullptr, s->h_enc);
        stream_rope_append_kernel<<<R, 256, 0, s->st>>>(s->qkv_enc, 3 * HQ, c.enc_heads, c.enc_head_dim, d_row_slot, d_row_pos,
                                                        ek + i * ring_stride, ev + i * ring_stride, ring, m->enc_cos, m->enc_sin);
        cuda_check(cudaGetLastError(), "stream_rope_append launch");
        launch_stream_attn(s->qkv_enc, R, 3 * HQ, c.enc_heads, c.enc_head_dim, d_row_slot, d_row_pos, ek + i * ring_stride,
                           ev + i * ring_stride, ring, c.enc_window, scale, s->attn_enc, s->st);
        s->linear(l.wo, s->attn_enc, R, s->x_enc, d, l.bo, s->x_enc, EPI_RESIDUAL);
        s->linear_n(l.w13, s->x_enc, R, s->act_enc, c.enc_ffn, nullptr, nullptr, EPI_SILU_MUL, l.ffn_norm, nullptr, s->h_enc);
        s->linear(l.w2, s->act_enc, R, s->x_enc, d, l.b2, s->x_enc, EPI_RESIDUAL);
    }
    launch_rmsnorm(s->x_enc, m->enc_norm, nullptr, s->h_enc, R, d, m->norm_eps, s->st);
}

void StreamPool::ensure_pages(Slot &sl, int positions) {
    const int need = (positions + KV_PAGE - 1) / KV_PAGE;
    VOX_CHECK(need <= s->kv_max_pages, VOX_ECAPACITY, "stream session needs %d decoder positions > capacity %d", positions, s->out_ld);
    while ((int)sl.pages.size() < need) {
        VOX_CHECK(!free_pages.empty(), VOX_ECAPACITY, "decoder KV page pool exhausted (%d pages)", s->kv_n_pages);
        sl.pages.push_back(free_pages.back());
        free_pages.pop_back();
    }
}

// rows[i] = slot id of batch row i: page tables, positions, fed-back tokens, audio pointers of this step
void StreamPool::upload_rows(const std::vector<int> &rows, bool with_tokens) {
    const vox_model_info &c = m->info;
    const int nb = (int)rows.size(), mp = s->kv_max_pages;
    std::vector<int> pt((size_t)nb * mp, 0), pos(nb), tok(nb), zero(nb, 0);
    std::vector<const float *> ar(nb);
    for (int i = 0; i < nb; ++i) {
        const Slot &sl = slots[rows[i]];
        for (size_t k = 0; k < sl.pages.size(); ++k) pt[(size_t)i * mp + k] = sl.pages[k];
        pos[i] = sl.pos;
        tok[i] = sl.last_tok;
        ar[i] = s->audio + ((size_t)rows[i] * s->S4_max + sl.pos) * c.dec_dim;
    }
    CUDA_OK(cudaMemcpyAsync(s->d_page_table, pt.data(), sizeof(int) * pt.size(), cudaMemcpyHostToDevice, s->st));
    CUDA_OK(cudaMemcpyAsync(s->d_pos, pos.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, s->st));
    CUDA_OK(cudaMemcpyAsync(s->d_outpos, zero.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, s->st));
    if (with_tokens) CUDA_OK(cudaMemcpyAsync(s->d_tok, tok.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, s->st));
    CUDA_OK(cudaMemcpyAsync(d_audio_rows, ar.data(), sizeof(float *) * nb, cudaMemcpyHostToDevice, s->st));
    CUDA_OK(cudaStreamSynchronize(s->st));  // the staging vectors die with this frame
}

void StreamPool::tick(vox_stream_stats *st_out) {
    const vox_model_info &c = m->info;
    CUDA_OK(cudaSetDevice(m->device));
    const int d = c.enc_dim, D = c.dec_dim, rf = c.reshape_factor, P = c.prefix_len;
    vox_stream_stats stats{};
    cudaEvent_t e0 = s->ev[0], e1 = s->ev[1];
    CUDA_OK(cudaEventRecord(e0, s->st));
    bool more = true;
    while (more) {
        more = false;
        // ---- front end: mel -> conv1 -> conv2 for every session, new encoder rows gathered into s->x_enc
        std::vector<int> row_slot, row_pos;
        struct Span { int slot, r0, n; };
        std::vector<Span> spans;
        for (int id = 0; id < max_sessions; ++id) {
            Slot &sl = slots[id];
            if (!sl.open || sl.drained) continue;
            int64_t tg[5];
            stream_progress(sl.n_samples, sl.ended, rf, P, tg);
            const int mel_t = (int)tg[0], c1_t = (int)tg[1];
            int enc_t = (int)tg[2];
            if (enc_t - sl.n_enc > max_new) {  // bounded by the ring slack: the rest in the next pass
                enc_t = sl.n_enc + max_new;
                more = true;
            }
            float *mel_s = s->mel_tm + (size_t)id * s->max_mel_frames * c.n_mels;
            float *c1_s = s->h1 + (size_t)id * s->T1_max * d;
            if (mel_t > sl.n_mel) {
                launch_mel(pcm + (size_t)id * cap_samples, 1, sl.n_samples, cap_samples, m->mel.window, m->mel.fb_vals, m->mel.fb_start,
                           m->mel.fb_len, m->mel.fb_stride, mel_s, mel_t, 0, s->st, sl.n_mel);
                stats.mel_frames += mel_t - sl.n_mel;
                sl.n_mel = mel_t;
            }
            if (c1_t > sl.n_c1) {
                launch_conv2_gemm(mel_s, m->conv1_w, m->conv1_b, c1_s + (size_t)sl.n_c1 * d, 1, sl.n_mel, c1_t - sl.n_c1, c.n_mels, d, s->st,
                                  sl.n_c1);
                sl.n_c1 = c1_t;
            }
            const int n_new = enc_t - sl.n_enc;
            if (n_new > 0) {
                const int r0 = (int)row_slot.size();
                launch_conv2_gemm(c1_s, m->conv2_w, m->conv2_b, s->x_enc + (size_t)r0 * d, 1, sl.n_c1, n_new, d, d, s->st, sl.n_enc);
                for (int i = 0; i < n_new; ++i) {
                    row_slot.push_back(id);
                    row_pos.push_back(sl.n_enc + i);
                }
                spans.push_back({id, r0, n_new});
            }
        }
        // ---- encoder layers over all new rows at once (KV rings), final norm, scatter to the sessions
        const int R = (int)row_slot.size();
        if (R > 0) {
            CUDA_OK(cudaMemcpyAsync(d_row_slot, row_slot.data(), sizeof(int) * R, cudaMemcpyHostToDevice, s->st));
            CUDA_OK(cudaMemcpyAsync(d_row_pos, row_pos.data(), sizeof(int) * R, cudaMemcpyHostToDevice, s->st));
            CUDA_OK(cudaStreamSynchronize(s->st));
            encoder_rows(R);
            stats.encoder_rows += R;
            for (const Span &sp : spans) {
                Slot &sl = slots[sp.slot];
                CUDA_OK(cudaMemcpyAsync(enc_out + ((size_t)sp.slot * s->S_max + sl.n_enc) * d, s->h_enc + (size_t)sp.r0 * d,
                                        sizeof(float) * (size_t)sp.n * d, cudaMemcpyDeviceToDevice, s->st));
                sl.n_enc += sp.n;
            }
        }
        // ---- x4 frame stack + adapter (adapter.rs:108-122, model.rs:745-749): 4 consecutive frames are contiguous
        for (int id = 0; id < max_sessions; ++id) {
            Slot &sl = slots[id];
            if (!sl.open || sl.drained) continue;
            const int emb_t = sl.n_enc / rf, n_new = emb_t - sl.n_emb;
            if (n_new <= 0) continue;
            const float *src = enc_out + ((size_t)id * s->S_max + (size_t)sl.n_emb * rf) * d;
            float *dst = s->audio + ((size_t)id * s->S4_max + sl.n_emb) * D;
            s->linear(m->adapter0, src, n_new, s->adapter_h, D, nullptr, nullptr, EPI_GELU);
            s->linear(m->adapter2, s->adapter_h, n_new, dst, D, nullptr, nullptr, EPI_NONE);
            sl.n_emb = emb_t;
        }
        // ---- decoder: prefill of sessions whose 38 prefix positions have their audio (model.rs:883-923)
        for (int id = 0; id < max_sessions; ++id) {
            Slot &sl = slots[id];
            if (!sl.open || sl.drained || sl.pos != 0 || sl.n_emb < P) continue;
            ensure_pages(sl, P + 1);
            upload_rows({id}, false);
            std::vector<int> prefix((size_t)P, 32);
            prefix[0] = 1;
            s->audio_rows_dev = nullptr;
            s->audio_base = s->audio + (size_t)id * s->S4_max * D;  // row 0 of the launch = this session
            s->prefill(1, P, prefix.data(), true);
            s->audio_base = nullptr;
            int tok = 0;
            CUDA_OK(cudaMemcpyAsync(&tok, s->d_tok, sizeof(int), cudaMemcpyDeviceToHost, s->st));
            CUDA_OK(cudaStreamSynchronize(s->st));
            sl.last_tok = tok;
            sl.ids.push_back(tok);
            sl.pos = P;  // cached positions; the next step is position P and consumes audio[P]
            stats.prefills += 1;
        }
        // ---- decoder steps: every session with a pending position shares one step (model.rs:938-960)
        for (;;) {
            std::vector<int> rows;
            for (int id = 0; id < max_sessions; ++id) {
                Slot &sl = slots[id];
                if (!sl.open || sl.drained || sl.pos < P) continue;
                // position p = sl.pos consumes audio[p]; the offline loop stops before the last embedding (model.rs:938)
                // (sl.pos = cached positions = index of the next position's audio embedding)
                const int last = sl.ended ? std::min(sl.n_emb - 1, final_enc(sl) / rf - 2) : sl.n_emb - 1;
                if (sl.pos <= last) rows.push_back(id);
            }
            if (rows.empty()) break;
            for (int id : rows) ensure_pages(slots[id], slots[id].pos + 1);
            upload_rows(rows, true);
            s->audio_rows_dev = d_audio_rows;
            s->decode_step((int)rows.size(), true);
            s->audio_rows_dev = nullptr;
            s->mega_steps_host += 1;
            std::vector<int> toks(rows.size());
            CUDA_OK(cudaMemcpyAsync(toks.data(), s->d_tok, sizeof(int) * rows.size(), cudaMemcpyDeviceToHost, s->st));
            CUDA_OK(cudaStreamSynchronize(s->st));
            for (size_t i = 0; i < rows.size(); ++i) {
                Slot &sl = slots[rows[i]];
                sl.last_tok = toks[i];
                sl.ids.push_back(toks[i]);
                sl.pos += 1;
            }
            stats.decode_steps += 1;
            stats.decode_rows += (int)rows.size();
        }
        for (int id = 0; id < max_sessions; ++id) {
            Slot &sl = slots[id];
            if (!sl.open || sl.drained || !sl.ended) continue;
            int64_t tg[5];
            stream_progress(sl.n_samples, true, rf, P, tg);
            if (sl.n_enc == (int)tg[2] && sl.n_emb == (int)tg[3] && (int64_t)sl.ids.size() == tg[4]) sl.drained = true;
        }
    }
    CUDA_OK(cudaEventRecord(e1, s->st));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.0f;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    stats.gpu_ms = ms;
    for (int id = 0; id < max_sessions; ++id)
        if (slots[id].open && !slots[id].drained) stats.live_sessions += 1;
    if (st_out) *st_out = stats;
}

// Q4VoxtralModel::encode_audio_with_cache (model.rs:790-799) = Q4AudioEncoder::forward_with_cache (437-452: conv stem on
// the chunk alone, layers extend the caches) + reshape_encoder_output + adapter.
int StreamPool::encode_chunk(int id, const float *mel, int T, float *out, size_t cap) {
    const vox_model_info &c = m->info;
    Slot &sl = slot(id);
    VOX_CHECK(sl.n_samples == pad_left(pad) && sl.n_audio == 0, VOX_EINVAL, "stream session %d is fed by push(): do not mix with encode_chunk", id);
    VOX_CHECK(T >= 1 && T <= s->max_mel_frames, VOX_EINVAL, "mel chunk of %d frames exceeds the pool's capacity %d", T, s->max_mel_frames);
    CUDA_OK(cudaSetDevice(m->device));
    const int d = c.enc_dim, D = c.dec_dim, rf = c.reshape_factor;
    const int T1 = conv_out(T), S = conv_out(T1), S4 = S / rf;
    VOX_CHECK(sl.n_enc + S <= m->enc_rope_len, VOX_ECAPACITY, "encoder positions %d exceed the RoPE table (%d)", sl.n_enc + S, m->enc_rope_len);
    VOX_CHECK(cap >= (size_t)S4 * D, VOX_ECAPACITY, "audio_embeds capacity %zu < %zu", cap, (size_t)S4 * D);
    CUDA_OK(cudaMemcpyAsync(s->mel, mel, sizeof(float) * (size_t)c.n_mels * T, cudaMemcpyHostToDevice, s->st));
    launch_transpose_mel(s->mel, s->mel_tm, 1, c.n_mels, T, s->st);
    launch_conv2_gemm(s->mel_tm, m->conv1_w, m->conv1_b, s->h1, 1, T, T1, c.n_mels, d, s->st);
    // rows beyond the ring slack go through the layers in several passes; the conv output of the whole chunk waits in
    // the session's (otherwise unused in chunk mode) encoder-output region
    float *conv = enc_out + (size_t)id * s->S_max * d;
    launch_conv2_gemm(s->h1, m->conv2_w, m->conv2_b, conv, 1, T1, S, d, d, s->st);
    for (int r0 = 0; r0 < S; r0 += max_new) {
        const int R = std::min(max_new, S - r0);
        std::vector<int> rs(R, id), rp(R);
        for (int i = 0; i < R; ++i) rp[i] = sl.n_enc + r0 + i;
        CUDA_OK(cudaMemcpyAsync(s->x_enc, conv + (size_t)r0 * d, sizeof(float) * (size_t)R * d, cudaMemcpyDeviceToDevice, s->st));
        CUDA_OK(cudaMemcpyAsync(d_row_slot, rs.data(), sizeof(int) * R, cudaMemcpyHostToDevice, s->st));
        CUDA_OK(cudaMemcpyAsync(d_row_pos, rp.data(), sizeof(int) * R, cudaMemcpyHostToDevice, s->st));
        CUDA_OK(cudaStreamSynchronize(s->st));
        encoder_rows(R);
        CUDA_OK(cudaMemcpyAsync(s->packed + (size_t)r0 * d, s->h_enc, sizeof(float) * (size_t)R * d, cudaMemcpyDeviceToDevice, s->st));
    }
    sl.n_enc += S;
    if (S4 > 0) {
        s->linear(m->adapter0, s->packed, S4, s->adapter_h, D, nullptr, nullptr, EPI_GELU);
        s->linear(m->adapter2, s->adapter_h, S4, s->audio, D, nullptr, nullptr, EPI_NONE);
        CUDA_OK(cudaMemcpyAsync(out, s->audio, sizeof(float) * (size_t)S4 * D, cudaMemcpyDeviceToHost, s->st));
    }
    CUDA_OK(cudaStreamSynchronize(s->st));
    return S4;
}

int StreamPool::final_enc(const Slot &sl) const {
    int64_t tg[5];
    stream_progress(sl.n_samples, true, m->info.reshape_factor, m->info.prefix_len, tg);
    return (int)tg[2];
}

const float *StreamPool::audio_embeds(int id, int *n) {
    Slot &sl = slot(id);
    *n = sl.n_emb;
    return s->audio + (size_t)id * s->S4_max * m->info.dec_dim;
}

}  // namespace vox
