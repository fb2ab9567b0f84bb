// audio_host.cpp -- host-side audio plumbing: peak_normalize (reference src/audio/io.rs:59-68),
// PadConfig/pad_audio (pad.rs:20-103), chunk plan (chunk.rs:122-166), TimeEmbedding
// (src/models/time_embedding.rs:41-71), and the constant tables of the mel front-end
// (Hann window mel.rs:345-349, Slaney filterbank mel.rs:260-339) that the GPU mel kernel consumes.
#include "audio_host.h"

#include <cmath>
#include <cstring>

#include "common.h"

namespace vox {

void peak_normalize(float *s, size_t n, float target) {
    float max_amp = 0.0f;
    for (size_t i = 0; i < n; ++i) max_amp = std::fmax(max_amp, std::fabs(s[i]));
    if (max_amp < 1e-10f) return;
    float scale = target / max_amp;
    for (size_t i = 0; i < n; ++i) s[i] *= scale;
}

void pad_config_default(vox_pad_config *c) {
    c->sample_rate = 16000;
    c->n_left_pad_tokens = 76;
    c->frame_rate = 12.5f;
    c->extra_right_pad_tokens = 17;
}

static size_t samples_per_token(const vox_pad_config &c) { return (size_t)((float)c.sample_rate / c.frame_rate); }

size_t pad_left(const vox_pad_config &c) { return (size_t)c.n_left_pad_tokens * samples_per_token(c); }

size_t pad_right(const vox_pad_config &c, size_t total) {
    size_t spt = samples_per_token(c);
    size_t rem = total % spt;
    size_t align = rem == 0 ? 0 : spt - rem;
    return align + (size_t)c.extra_right_pad_tokens * spt;
}

size_t pad_audio_len(size_t n, const vox_pad_config &c) {
    size_t left = pad_left(c);
    return left + n + pad_right(c, n + left);
}

std::vector<vox_chunk> chunk_plan(size_t n, size_t max_mel_frames, size_t overlap_frames) {
    const size_t hop = 160;
    std::vector<vox_chunk> out;
    size_t pos = 0, idx = 0;
    size_t max_chunk = max_mel_frames * hop;
    size_t step = (max_mel_frames - overlap_frames) * hop;
    VOX_CHECK(max_mel_frames > overlap_frames, VOX_EINVAL, "overlap_frames must be < max_mel_frames");
    while (pos < n) {
        size_t end = pos + max_chunk < n ? pos + max_chunk : n;
        vox_chunk c;
        c.start_sample = pos;
        c.end_sample = end;
        c.index = idx++;
        c.is_last = end >= n;
        out.push_back(c);
        pos += step;
    }
    return out;
}

void time_embedding(float t, int dim, float *out) {
    int half = dim / 2;
    float log_theta = std::log(10000.0f);
    for (int i = 0; i < half; ++i) {
        float freq = std::exp(-log_theta * (float)i / (float)half);
        float ang = t * freq;
        out[i] = std::cos(ang);
        out[half + i] = std::sin(ang);
    }
}

void hann_window(int n, float *out) {
    const float pi = 3.14159265358979323846f;
    for (int i = 0; i < n; ++i) out[i] = 0.5f * (1.0f - std::cos(2.0f * pi * (float)i / (float)n));
}

static float hz_to_mel(float f) {
    const float f_sp = 200.0f / 3.0f, min_log_hz = 1000.0f, min_log_mel = min_log_hz / f_sp;
    const float logstep = 0.06875174f;
    return f < min_log_hz ? f / f_sp : min_log_mel + std::log(f / min_log_hz) / logstep;
}

static float mel_to_hz(float m) {
    const float f_sp = 200.0f / 3.0f, min_log_hz = 1000.0f, min_log_mel = min_log_hz / f_sp;
    const float logstep = 0.06875174f;
    return m < min_log_mel ? m * f_sp : min_log_hz * std::exp((m - min_log_mel) * logstep);
}

void mel_filterbank(float *fb /* [128][201] */) {
    const int n_mels = kMelBins, n_freq = kMelFreqs, n_fft = kMelNfft;
    const float sr = 16000.0f;
    std::memset(fb, 0, sizeof(float) * n_mels * n_freq);
    float mel_min = hz_to_mel(0.0f), mel_max = hz_to_mel(sr / 2.0f);
    std::vector<float> hz(n_mels + 2);
    for (int i = 0; i < n_mels + 2; ++i) hz[i] = mel_to_hz(mel_min + (mel_max - mel_min) * (float)i / (float)(n_mels + 1));
    for (int i = 0; i < n_mels; ++i) {
        float lo = hz[i], ce = hz[i + 1], up = hz[i + 2];
        float *row = fb + (size_t)i * n_freq;
        for (int j = 0; j < n_freq; ++j) {
            float fr = (float)j * sr / (float)n_fft;
            if (fr >= lo && fr <= ce && ce > lo) row[j] = (fr - lo) / (ce - lo);
            else if (fr > ce && fr <= up && up > ce) row[j] = (up - fr) / (up - ce);
        }
        float bw = hz[i + 2] - hz[i];
        if (bw > 0.0f) {
            float e = 2.0f / bw;
            for (int j = 0; j < n_freq; ++j) row[j] *= e;
        }
    }
}

// Stage i's output t is final when the inputs it reads are: an STFT frame needs samples [160 t - 200, 160 t + 200)
// (centre padding; the reflect pad only ever sees the zeros pad_audio adds), a k3 s2 p1 convolution output needs input
// frame 2 t + 1, an audio embedding needs `reshape_factor` encoder frames, decoder position p needs embedding p - 1
// (the prefill needs the first `prefix_len`), and the last embedding of a finished stream is never consumed
// (model.rs:938: the loop stops at S - 1).
void stream_progress(size_t n_samples, bool ended, int reshape_factor, int prefix_len, int64_t out[5]) {
    const int64_t n = (int64_t)n_samples;
    auto conv_final = [&](int64_t t_in) -> int64_t {
        if (t_in <= 0) return 0;
        if (ended) return (t_in + 2 - 3) / 2 + 1;   // conv.rs:47-48
        return t_in >= 2 ? (t_in - 2) / 2 + 1 : 0;   // largest t with 2 t + 1 <= t_in - 1
    };
    const int64_t mel = ended ? n / 160 : (n >= 200 ? (n - 200) / 160 + 1 : 0);
    const int64_t c1 = conv_final(mel);
    const int64_t enc = conv_final(c1);
    const int64_t emb = reshape_factor > 0 ? enc / reshape_factor : 0;
    int64_t ids = 0;
    if (emb >= prefix_len) {
        const int64_t last_pos = ended ? emb - 1 : emb;
        ids = 1 + (last_pos > prefix_len ? last_pos - prefix_len : 0);
    }
    out[0] = mel;
    out[1] = c1;
    out[2] = enc;
    out[3] = emb;
    out[4] = ids;
}

}  // namespace vox
