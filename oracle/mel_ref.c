/* oracle/mel_ref.c -- scalar single-threaded CPU restatement of MelSpectrogram::compute_log.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): checker + timed CPU baseline ("port").
 *
 * Follows src/audio/mel.rs:
 *   hann_window 345-349, create_mel_filterbank 288-339 (hz_to_mel/mel_to_hz 260-285),
 *   stft 185-244 (reflect pad 200, frames i*160, last frame dropped), power (110-114),
 *   apply_mel_filterbank 247-257 (sequential f32 sum over 201 bins),
 *   compute_log 128-165 (log10(max(.,1e-10)), max(., 1.5-8), (x+4)/4).
 * The reference's FFT is rustfft (absent third-party crate); here a plain recursive
 * mixed-radix (2,5) decimation-in-time FFT in f32 -- same O(n log n) work per frame,
 * like the reference re-planned per call and run per frame on one thread.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define N_FFT 400
#define HOP 160
#define N_MELS 128
#define N_FREQ 201

typedef struct { float re, im; } cpx;

static void fft_rec(const cpx *in, cpx *out, int n, int stride, const cpx *tw, int tws) {
    if (n == 1) { out[0] = in[0]; return; }
    int p = (n % 2 == 0) ? 2 : 5;
    int m = n / p;
    for (int r = 0; r < p; ++r) fft_rec(in + r * stride, out + r * m, m, stride * p, tw, tws * p);
    cpx tmp[5];
    for (int k = 0; k < m; ++k) {
        for (int q = 0; q < p; ++q) {
            float sr = 0.f, si = 0.f;
            int kk = k + q * m;
            for (int r = 0; r < p; ++r) {
                cpx w = tw[((long)r * kk * tws) % N_FFT];
                cpx v = out[r * m + k];
                sr += v.re * w.re - v.im * w.im;
                si += v.re * w.im + v.im * w.re;
            }
            tmp[q].re = sr; tmp[q].im = si;
        }
        for (int q = 0; q < p; ++q) out[k + q * m] = tmp[q];
    }
}

static float hz_to_mel(float f) {
    const float F_SP = 200.0f / 3.0f, MIN_LOG_HZ = 1000.0f, MIN_LOG_MEL = MIN_LOG_HZ / F_SP;
    const float LOGSTEP = 0.06875174f;
    return f < MIN_LOG_HZ ? f / F_SP : MIN_LOG_MEL + logf(f / MIN_LOG_HZ) / LOGSTEP;
}
static float mel_to_hz(float m) {
    const float F_SP = 200.0f / 3.0f, MIN_LOG_HZ = 1000.0f, MIN_LOG_MEL = MIN_LOG_HZ / F_SP;
    const float LOGSTEP = 0.06875174f;
    return m < MIN_LOG_MEL ? m * F_SP : MIN_LOG_HZ * expf((m - MIN_LOG_MEL) * LOGSTEP);
}

/* samples[n] -> out[frames*128] (row-major [frame][mel]); returns frames. */
long long oracle_mel_compute_log(const float *samples, long long n, float *out) {
    static float fb[N_MELS][N_FREQ];
    float window[N_FFT];
    cpx tw[N_FFT];
    /* per-call setup, like MelSpectrogram::new + the per-call FftPlanner (mel.rs:73-91,208) */
    for (int i = 0; i < N_FFT; ++i) {
        window[i] = 0.5f * (1.0f - cosf(2.0f * (float)M_PI * (float)i / (float)N_FFT));
        double a = -2.0 * M_PI * (double)i / (double)N_FFT;
        tw[i].re = (float)cos(a); tw[i].im = (float)sin(a);
    }
    float hz[N_MELS + 2];
    float mel_min = hz_to_mel(0.0f), mel_max = hz_to_mel(8000.0f);
    for (int i = 0; i < N_MELS + 2; ++i)
        hz[i] = mel_to_hz(mel_min + (mel_max - mel_min) * (float)i / (float)(N_MELS + 1));
    memset(fb, 0, sizeof(fb));
    for (int i = 0; i < N_MELS; ++i) {
        float lo = hz[i], ce = hz[i + 1], up = hz[i + 2];
        for (int j = 0; j < N_FREQ; ++j) {
            float fr = (float)j * 16000.0f / (float)N_FFT;
            if (fr >= lo && fr <= ce && ce > lo) fb[i][j] = (fr - lo) / (ce - lo);
            else if (fr > ce && fr <= up && up > ce) fb[i][j] = (up - fr) / (up - ce);
        }
        float bw = hz[i + 2] - hz[i];
        if (bw > 0.0f) { float e = 2.0f / bw; for (int j = 0; j < N_FREQ; ++j) fb[i][j] *= e; }
    }
    const int pad = N_FFT / 2;
    long long plen = n + 2 * pad;
    float *padded = (float *)malloc(sizeof(float) * (size_t)plen);
    for (int i = pad, o = 0; i >= 1; --i, ++o) {
        long long idx = i < (n > 0 ? n - 1 : 0) ? i : (n > 0 ? n - 1 : 0);
        padded[o] = (idx < n) ? samples[idx] : 0.0f;
    }
    memcpy(padded + pad, samples, sizeof(float) * (size_t)n);
    for (int i = 0; i < pad; ++i) {
        long long idx = (n >= 2 ? n - 2 : 0) - i;
        if (idx < 0) idx = 0;
        padded[pad + n + i] = (idx < n) ? samples[idx] : 0.0f;
    }
    long long frames = (plen - N_FFT) / HOP;
    if (frames < 0) frames = 0;
    const float min_val = 1.5f - 8.0f;
    cpx buf[N_FFT], spec[N_FFT];
    float power[N_FREQ];
    for (long long f = 0; f < frames; ++f) {
        const float *src = padded + f * HOP;
        for (int j = 0; j < N_FFT; ++j) { buf[j].re = src[j] * window[j]; buf[j].im = 0.0f; }
        fft_rec(buf, spec, N_FFT, 1, tw, 1);
        for (int j = 0; j < N_FREQ; ++j) power[j] = spec[j].re * spec[j].re + spec[j].im * spec[j].im;
        for (int m = 0; m < N_MELS; ++m) {
            float acc = 0.0f;
            for (int j = 0; j < N_FREQ; ++j) acc += fb[m][j] * power[j];
            float v = log10f(fmaxf(acc, 1e-10f));
            v = fmaxf(v, min_val);
            out[f * N_MELS + m] = (v + 4.0f) / 4.0f;
        }
    }
    free(padded);
    return frames;
}
