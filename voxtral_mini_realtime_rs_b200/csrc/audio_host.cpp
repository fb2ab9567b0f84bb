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

}  // namespace vox
