"""Audio front-end oracle (test infrastructure, see oracle/__init__.py).

Restates in numpy f32:
  * AudioBuffer::peak_normalize         src/audio/io.rs:59-68
  * PadConfig / pad_audio               src/audio/pad.rs:20-103
  * ChunkConfig / chunk_audio           src/audio/chunk.rs:9-166
  * MelSpectrogram (Hann, Slaney filterbank, STFT, log-mel)   src/audio/mel.rs:73-349
  * TimeEmbedding::embed                src/models/time_embedding.rs:41-71

The reference's FFT is rustfft 6.4 (third-party, absent); here it is numpy's f32 pocketfft.
Both are O(eps_f32 * log n) accurate but round differently, and the reference's golden
`.npy` mel fixtures (mel.rs:486-614) are absent => **parity unpinned for the FFT rounding**.
Pinned closed-form cases (tests/test_oracle_pins.py): Hann(400)[1] (mel.rs:384-396),
Hann(4) (398-406), 1 s silence (409-421), 440 Hz sine range (438-456), frame counts
(459-465), hz<->mel round trips (468-483), pad sample counts (pad.rs:139-218),
time embedding dim-4 (time_embedding.rs:91-128), chunk plans (chunk.rs tests).
"""
from __future__ import annotations

import math
import numpy as np

F32 = np.float32

SAMPLE_RATE = 16000
N_FFT = 400
HOP = 160
N_MELS = 128
LOG_MEL_MAX = F32(1.5)


# ---------------------------------------------------------------- io.rs:59-68
def peak_normalize(samples: np.ndarray, target_peak: float = 0.95) -> np.ndarray:
    s = np.asarray(samples, F32).copy()
    max_amp = F32(np.max(np.abs(s))) if s.size else F32(0)
    if max_amp < F32(1e-10):
        return s
    scale = F32(target_peak) / max_amp
    return (s * scale).astype(F32)


# ---------------------------------------------------------------- pad.rs:20-103
class PadConfig:
    def __init__(self, sample_rate=16000, n_left_pad_tokens=76, frame_rate=12.5,
                 extra_right_pad_tokens=17):
        self.sample_rate = sample_rate
        self.n_left_pad_tokens = n_left_pad_tokens
        self.frame_rate = frame_rate
        self.extra_right_pad_tokens = extra_right_pad_tokens

    def samples_per_token(self) -> int:
        return int(F32(self.sample_rate) / F32(self.frame_rate))

    def left_pad_samples(self) -> int:
        return self.n_left_pad_tokens * self.samples_per_token()

    def right_pad_samples(self, total_samples: int) -> int:
        spt = self.samples_per_token()
        rem = total_samples % spt
        align = 0 if rem == 0 else spt - rem
        return align + self.extra_right_pad_tokens * spt


def pad_audio(samples: np.ndarray, cfg: PadConfig | None = None) -> np.ndarray:
    cfg = cfg or PadConfig()
    s = np.asarray(samples, F32)
    left = cfg.left_pad_samples()
    right = cfg.right_pad_samples(s.size + left)
    out = np.zeros(left + s.size + right, F32)
    out[left:left + s.size] = s
    return out


def num_audio_tokens(samples: int, cfg: PadConfig | None = None) -> int:
    cfg = cfg or PadConfig()
    return samples // cfg.samples_per_token()


# ---------------------------------------------------------------- chunk.rs:9-166
def chunk_plan(n_samples: int, max_mel_frames: int = 1500, hop: int = HOP, overlap_frames: int = 0):
    """Returns [(start, end, index, is_last)] exactly as ChunkIterator (chunk.rs:122-150)."""
    out = []
    pos, idx = 0, 0
    max_chunk = max_mel_frames * hop
    step = (max_mel_frames - overlap_frames) * hop
    while pos < n_samples:
        end = min(pos + max_chunk, n_samples)
        out.append((pos, end, idx, end >= n_samples))
        pos += step
        idx += 1
    return out


def needs_chunking(n_samples: int, max_mel_frames: int = 1500, hop: int = HOP) -> bool:
    return n_samples > max_mel_frames * hop


# ---------------------------------------------------------------- mel.rs:260-349
_F_SP = F32(200.0) / F32(3.0)
_MIN_LOG_HZ = F32(1000.0)
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = F32(0.068_751_74)


def hz_to_mel(f) -> np.float32:
    f = F32(f)
    if f < _MIN_LOG_HZ:
        return F32(f / _F_SP)
    return F32(_MIN_LOG_MEL + F32(np.log(F32(f / _MIN_LOG_HZ))) / _LOGSTEP)


def mel_to_hz(m) -> np.float32:
    m = F32(m)
    if m < _MIN_LOG_MEL:
        return F32(m * _F_SP)
    return F32(_MIN_LOG_HZ * F32(np.exp(F32(F32(m - _MIN_LOG_MEL) * _LOGSTEP))))


def hann_window(length: int) -> np.ndarray:
    i = np.arange(length, dtype=F32)
    two_pi = F32(2.0) * F32(math.pi)
    return (F32(0.5) * (F32(1.0) - np.cos((two_pi * i / F32(length)).astype(F32)).astype(F32))).astype(F32)


def create_mel_filterbank(sample_rate=SAMPLE_RATE, n_fft=N_FFT, n_mels=N_MELS, fmin=0.0,
                          fmax=None) -> np.ndarray:
    """mel.rs:288-339 -> float32 [n_mels, n_fft/2+1]."""
    if fmax is None:
        fmax = F32(sample_rate) / F32(2.0)
    n_freqs = n_fft // 2 + 1
    mel_min, mel_max = hz_to_mel(fmin), hz_to_mel(fmax)
    mel_points = [F32(mel_min + F32(F32(mel_max - mel_min) * F32(i)) / F32(n_mels + 1))
                  for i in range(n_mels + 2)]
    hz = [mel_to_hz(m) for m in mel_points]
    freqs = [F32(F32(j) * F32(sample_rate) / F32(n_fft)) for j in range(n_freqs)]
    fb = np.zeros((n_mels, n_freqs), F32)
    for i in range(n_mels):
        lo, ce, up = hz[i], hz[i + 1], hz[i + 2]
        for j, fr in enumerate(freqs):
            if lo <= fr <= ce and ce > lo:
                fb[i, j] = F32(fr - lo) / F32(ce - lo)
            elif ce < fr <= up and up > ce:
                fb[i, j] = F32(up - fr) / F32(up - ce)
        bw = F32(hz[i + 2] - hz[i])
        if bw > 0:
            fb[i] = (fb[i] * (F32(2.0) / bw)).astype(F32)
    return fb


def num_frames(num_samples: int) -> int:
    """mel.rs:175-182."""
    return (num_samples + 2 * (N_FFT // 2) - N_FFT) // HOP


def reflect_pad(samples: np.ndarray, pad: int = N_FFT // 2) -> np.ndarray:
    """mel.rs:190-205 (torch.stft center=True reflect)."""
    s = np.asarray(samples, F32)
    n = s.size
    left = [s[min(i, max(n - 1, 0))] if n else F32(0) for i in range(pad, 0, -1)]
    right = []
    for i in range(pad):
        idx = max(max(n - 2, 0) - i, 0)
        right.append(s[idx] if idx < n else F32(0))
    return np.concatenate([np.asarray(left, F32), s, np.asarray(right, F32)]).astype(F32)


class MelSpectrogram:
    """MelSpectrogram::voxtral() (mel.rs:73-100)."""

    def __init__(self):
        self.mel_basis = create_mel_filterbank()
        self.window = hann_window(N_FFT)

    def power_frames(self, samples: np.ndarray) -> np.ndarray:
        padded = reflect_pad(samples)
        n_frames = (padded.size - N_FFT) // HOP
        if n_frames <= 0:
            return np.zeros((0, N_FFT // 2 + 1), F32)
        idx = np.arange(n_frames)[:, None] * HOP + np.arange(N_FFT)[None, :]
        frames = (padded[idx] * self.window[None, :]).astype(F32)
        spec = np.fft.rfft(frames, axis=1)  # numpy>=2: single-precision FFT for f32 input
        re = spec.real.astype(F32)
        im = spec.imag.astype(F32)
        return (re * re + im * im).astype(F32)

    def compute(self, samples: np.ndarray) -> np.ndarray:
        p = self.power_frames(samples)
        acc = np.zeros((p.shape[0], N_MELS), F32)
        # sequential f32 sum over the 201 bins (Rust `.sum()` fold), mel.rs:247-257
        for j in range(p.shape[1]):
            acc += (self.mel_basis[None, :, j] * p[:, j:j + 1]).astype(F32)
        return acc

    def compute_log(self, samples: np.ndarray) -> np.ndarray:
        """mel.rs:128-165 -> float32 [n_frames, 128]."""
        mel = self.compute(samples)
        log_mel = np.log10(np.maximum(mel, F32(1e-10))).astype(F32)
        min_val = F32(LOG_MEL_MAX - F32(8.0))
        log_mel = np.maximum(log_mel, min_val)
        return ((log_mel + F32(4.0)) / F32(4.0)).astype(F32)


def mel_tensor_from_audio(samples: np.ndarray) -> np.ndarray:
    """transcribe.rs:279-306: pad -> compute_log -> transpose -> [1,128,T]."""
    mel = MelSpectrogram().compute_log(pad_audio(samples))
    return np.ascontiguousarray(mel.T)[None]


# ---------------------------------------------------------------- time_embedding.rs:41-71
def time_embedding(t: float, dim: int = 3072, theta: float = 10000.0) -> np.ndarray:
    half = dim // 2
    log_theta = F32(np.log(F32(theta)))
    i = np.arange(half, dtype=F32)
    inv_freq = np.exp((-log_theta * i / F32(half)).astype(F32)).astype(F32)
    ang = (F32(t) * inv_freq).astype(F32)
    return np.concatenate([np.cos(ang), np.sin(ang)]).astype(F32)


# ---------------------------------------------------------------- synthetic signals
# pure data generators, shared with bench.py
from voxtral_mini_realtime_rs_b200.synth import noise_chirp, sine_16k, speechlike  # noqa: E402,F401
