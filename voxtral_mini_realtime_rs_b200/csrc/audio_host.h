#pragma once
#include <cstddef>
#include <vector>

#include "../../include/voxtral.h"

namespace vox {

constexpr int kMelBins = 128;
constexpr int kMelFreqs = 201;
constexpr int kMelNfft = 400;
constexpr int kMelHop = 160;

void peak_normalize(float *s, size_t n, float target);
void pad_config_default(vox_pad_config *c);
size_t pad_left(const vox_pad_config &c);
size_t pad_right(const vox_pad_config &c, size_t total);
size_t pad_audio_len(size_t n, const vox_pad_config &c);
std::vector<vox_chunk> chunk_plan(size_t n, size_t max_mel_frames, size_t overlap_frames);
// incremental stage counts for n known samples of the padded signal (see include/voxtral.h vox_stream_progress)
void stream_progress(size_t n_samples, bool ended, int reshape_factor, int prefix_len, int64_t out[5]);
void time_embedding(float t, int dim, float *out);
void hann_window(int n, float *out);
void mel_filterbank(float *fb);
inline size_t mel_num_frames(size_t n) { return (n + 2 * (kMelNfft / 2) - kMelNfft) / kMelHop; }

}  // namespace vox
