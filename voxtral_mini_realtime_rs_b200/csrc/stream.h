// stream.h -- pool of live streaming sessions sharing one GPU worker (stream.cu).
#pragma once
#include <cstdint>
#include <vector>

#include "model.h"

namespace vox {

struct StreamPool {
    struct Slot {
        bool open = false, ended = false, drained = false;
        size_t n_samples = 0, n_audio = 0;        // padded samples known so far / audio samples pushed
        int n_mel = 0, n_c1 = 0, n_enc = 0, n_emb = 0;  // final frames produced per stage
        int pos = 0;                              // decoder positions cached (0: prefill pending)
        int last_tok = 0;
        std::vector<int32_t> ids;                 // emitted ids (positions >= 38)
        size_t polled = 0;
        std::vector<int> pages;                   // decoder KV pages owned (logical order)
    };
    Model *m = nullptr;
    Session *s = nullptr;       // private session: weights view, decoder state, workspaces, stream
    vox_pad_config pad{};
    int max_sessions = 0, max_new = 0, ring = 0;
    size_t cap_samples = 0;
    float *pcm = nullptr;       // [slot][cap_samples] padded signal
    float *enc_out = nullptr;   // [slot][S_max][enc_dim] encoder output frames (after the final norm)
    float *ek = nullptr, *ev = nullptr;  // encoder K/V rings [layer][slot][ring][H*hd], absolute position p at p % ring
    int *d_row_slot = nullptr, *d_row_pos = nullptr;
    const float **d_audio_rows = nullptr;
    std::vector<Slot> slots;
    std::vector<int> free_pages;

    static StreamPool *create(Model *m, int max_sessions, float max_seconds);
    ~StreamPool();
    int open();
    void push(int id, const float *samples, size_t n);
    void finish(int id);
    void close(int id);
    void tick(vox_stream_stats *stats);
    size_t poll(int id, int32_t *ids, size_t cap, bool *done);
    const float *audio_embeds(int id, int *n);   // device pointer [n][dec_dim]
    // encode_audio_with_cache (model.rs:790-799): one mel chunk [128][T] (host) through the conv stem ON ITS OWN (zero
    // padding at the chunk edges, as upstream) and the encoder layers over the session's K/V rings; returns the
    // chunk's S/4 audio embeddings (host, [n][dec_dim]).  For sessions driven chunk-wise instead of push()/tick().
    int encode_chunk(int id, const float *mel, int T, float *out, size_t cap_floats);

  private:
    Slot &slot(int id);
    void encoder_rows(int R);
    void ensure_pages(Slot &sl, int positions);
    void upload_rows(const std::vector<int> &rows, bool with_tokens);
    int final_enc(const Slot &sl) const;
};

}  // namespace vox
