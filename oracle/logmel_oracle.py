"""CPU oracle for the Audio-CALM log-mel front-end and its statistics pass.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may call it, and only as the checker (or as the timed CPU baseline), never as the thing shipped.

It is a numpy restatement (fp64 "truth" mode and an fp32 mode that follows the reference's operation
order) of what the reference computes with torchaudio / torch.stft.  Every function cites the reference
lines it follows (paths relative to the reference repository root).

Parity pinning: the reference ships no tests or known-answer vectors for this path, so
the oracle is pinned against outputs of the reference itself, generated in the build container by
``oracle/gen_golden.py`` (imports ``preprocess/core.py`` from the read-only reference checkout) and
committed under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks this module against every
one of those vectors.
"""
from __future__ import annotations

import math
from typing import Iterable, Optional, Sequence, Tuple

import numpy as np

# ---------------------------------------------------------------------------------------------
# constants of the "calm" preset: preprocess/core.py:33 (constructor defaults) and :37-48
# ---------------------------------------------------------------------------------------------
SAMPLE_RATE = 16000
N_FFT = 1024
HOP = 256
N_MELS = 80
F_MIN = 0.0
F_MAX = 8000.0
CLAMP_MIN = 1e-5                      # preprocess/core.py:60
LOG_FLOOR = math.log(1e-5)            # -11.512925...
MEL_MEAN_DEFAULT = -6.589515          # models/modeling_vae.py:317
MEL_STD_DEFAULT = 3.860679            # models/modeling_vae.py:318


# ---------------------------------------------------------------------------------------------
# deterministic synthetic inputs (RNG independent)
# ---------------------------------------------------------------------------------------------
def hash_noise(n: int, seed: int) -> np.ndarray:
    """splitmix64-style hash noise in [-0.5, 0.5), float32, identical on every platform."""
    i = np.arange(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = i + (np.uint64(seed) << np.uint64(32)) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)).astype(np.float64) / 2.0 ** 24 - 0.5).astype(np.float32)


def synth_clip(n: int, seed: int) -> np.ndarray:
    """Speech-like synthetic clip used by the parity tests: hash noise with a slow envelope
    ``0.25 + 0.75 sin^2(2 pi 0.7 t)`` and the final 5 % exactly zero (exercises the clamp)."""
    t = np.arange(n, dtype=np.float64) / SAMPLE_RATE
    env = 0.25 + 0.75 * np.sin(2.0 * np.pi * 0.7 * t) ** 2
    x = (hash_noise(n, seed).astype(np.float64) * 0.4 * env).astype(np.float32)
    x[n - n // 20:] = 0.0
    return x


def bench_clip(n: int, seed: int) -> np.ndarray:
    """A clip of the BENCHMARK's distribution (bench.synth_batch: N(0, 0.1^2) noise under the slow envelope, clipped to +-1, last 5 %
    exact zeros) made RNG-independent: the Gaussian is the sum of four hash-noise uniforms scaled to sigma = 0.1."""
    t = np.arange(n, dtype=np.float64) / SAMPLE_RATE
    env = 0.25 + 0.75 * np.sin(2.0 * np.pi * 0.7 * t) ** 2
    g = sum(hash_noise(n, seed + 1000 * k).astype(np.float64) for k in range(4)) * (0.1 / np.sqrt(4.0 / 12.0))
    x = np.clip(g * env, -1.0, 1.0).astype(np.float32)
    x[n - n // 20:] = 0.0
    return x


# ---------------------------------------------------------------------------------------------
# tables (fp64 derivations; the product builds its fp32 tables with torch so that they are
# bit-identical to torchaudio's -- tests compare both)
# ---------------------------------------------------------------------------------------------
def hann_window_f64(n: int = N_FFT) -> np.ndarray:
    """Periodic Hann window, as ``torch.hann_window(n)`` (torchaudio Spectrogram default used at
    preprocess/core.py:37-48; win_length defaults to n_fft)."""
    k = np.arange(n, dtype=np.float64)
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * k / n)


def _hz_to_mel_slaney(f: float) -> float:
    f_sp = 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    if f >= min_log_hz:
        return min_log_mel + math.log(f / min_log_hz) / logstep
    return f / f_sp


def _mel_to_hz_slaney(m: np.ndarray) -> np.ndarray:
    f_sp = 200.0 / 3.0
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    f = f_sp * m
    log_t = m >= min_log_mel
    f[log_t] = min_log_hz * np.exp(logstep * (m[log_t] - min_log_mel))
    return f


def slaney_fbanks_f64(n_freqs: int = N_FFT // 2 + 1, f_min: float = F_MIN, f_max: float = F_MAX,
                      n_mels: int = N_MELS, sample_rate: int = SAMPLE_RATE) -> np.ndarray:
    """Slaney-scale, slaney-normalised triangular filterbank ``[n_freqs, n_mels]`` -- the algorithm
    behind ``melscale_fbanks(513, 0, 8000, 80, 16000, "slaney", "slaney")`` that
    ``MelSpectrogram(norm="slaney", mel_scale="slaney")`` builds at preprocess/core.py:37-48."""
    all_freqs = np.linspace(0.0, sample_rate // 2, n_freqs)
    m_pts = np.linspace(_hz_to_mel_slaney(f_min), _hz_to_mel_slaney(f_max), n_mels + 2)
    f_pts = _mel_to_hz_slaney(m_pts.copy())
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
    return fb * enorm[None, :]


# ---------------------------------------------------------------------------------------------
# frame bookkeeping (must be bit-exact)
# ---------------------------------------------------------------------------------------------
def frames_for_length(length: int, hop: int = HOP, n_fft: int = N_FFT) -> int:
    """``T = 1 + L // hop`` for ``torch.stft(center=True)`` (preprocess/core.py:55).  Reflect padding by
    ``n_fft // 2`` needs ``L > n_fft // 2``; torch raises RuntimeError otherwise."""
    if length <= n_fft // 2:
        raise RuntimeError(
            f"reflect padding of {n_fft // 2} needs an input longer than {n_fft // 2} samples, got {length}")
    return 1 + length // hop


def padded_frames(t: int, multiple: int = 4) -> int:
    """Frame count after the pad-to-4 step (preprocess/process_dataset.py:146-150)."""
    return t if t % multiple == 0 else t + (multiple - t % multiple)


# ---------------------------------------------------------------------------------------------
# a1: process_audio_chunk -- preprocess/core.py:93-112
# ---------------------------------------------------------------------------------------------
def process_audio_chunk(wav: np.ndarray) -> np.ndarray:
    """``[C, L]`` float32 -> ``[1, L]`` float32.  Channel mean if C > 1 (core.py:102-103); then, if the
    peak is > 0, ``wav / (peak + 1e-8) * 0.95`` in float32 with that operation order (core.py:108-110)."""
    wav = np.asarray(wav, dtype=np.float32)
    if wav.ndim != 2:
        raise ValueError("expected [C, L]")
    if wav.shape[0] > 1:
        # torch.mean over dim 0 in fp32: sum of C values then divide
        acc = np.zeros(wav.shape[1], dtype=np.float32)
        for c in range(wav.shape[0]):
            acc = (acc + wav[c]).astype(np.float32)
        wav = (acc / np.float32(wav.shape[0])).astype(np.float32)[None, :]
    peak = np.float32(np.max(np.abs(wav))) if wav.size else np.float32(0.0)
    if peak > 0:
        denom = np.float32(peak + np.float32(1e-8))
        wav = ((wav / denom).astype(np.float32) * np.float32(0.95)).astype(np.float32)
    return wav


# ---------------------------------------------------------------------------------------------
# a3-a5: MelExtractor.forward -- preprocess/core.py:50-61
# ---------------------------------------------------------------------------------------------
def reflect_pad(x: np.ndarray, pad: int) -> np.ndarray:
    """``F.pad(x, (pad, pad), mode="reflect")`` on the last axis (torch.stft center=True)."""
    if x.shape[-1] <= pad:
        raise RuntimeError("reflect padding needs an input longer than the pad")
    return np.concatenate([x[..., pad:0:-1], x, x[..., -2:-pad - 2:-1]], axis=-1)


def power_spectrogram(wav: np.ndarray, window: np.ndarray, hop: int = HOP, dtype=np.float64) -> np.ndarray:
    """``[L]`` -> ``[n_fft//2+1, T]`` power spectrogram: reflect-pad n_fft//2, frame at ``hop``, multiply
    by ``window``, one-sided DFT, ``abs() ** 2`` (torchaudio.functional.spectrogram with power=2.0,
    normalized=False, as configured at preprocess/core.py:37-48)."""
    n_fft = window.shape[0]
    x = reflect_pad(np.asarray(wav, dtype=dtype), n_fft // 2)
    t = 1 + (x.shape[0] - n_fft) // hop
    idx = np.arange(n_fft)[None, :] + hop * np.arange(t)[:, None]
    frames = x[idx] * window.astype(dtype)[None, :]
    spec = np.fft.rfft(frames.astype(np.float64), axis=-1)
    if dtype == np.float32:
        spec = spec.astype(np.complex64)
        mag = np.abs(spec).astype(np.float32)          # abs() then pow(2.0), as torchaudio does
        return (mag * mag).T.astype(np.float32)
    return (spec.real ** 2 + spec.imag ** 2).T


def logmel(wav: np.ndarray, window: Optional[np.ndarray] = None, fb: Optional[np.ndarray] = None,
           hop: int = HOP, clamp_min: float = CLAMP_MIN, dtype=np.float64) -> np.ndarray:
    """``[L]`` -> ``[n_mels, 1 + L//hop]``: ``log(clamp(fb^T @ |STFT|^2, min=1e-5))`` (preprocess/core.py:55,60).
    ``window`` / ``fb`` default to the fp64-derived tables; pass the product's fp32 tables to isolate
    kernel arithmetic from table derivation."""
    if window is None:
        window = hann_window_f64(N_FFT)
    if fb is None:
        fb = slaney_fbanks_f64()
    p = power_spectrogram(wav, window, hop=hop, dtype=dtype)          # [F, T]
    mel = fb.astype(dtype).T @ p                                       # [n_mels, T]
    return np.log(np.maximum(mel, dtype(clamp_min))).astype(dtype)


# ---------------------------------------------------------------------------------------------
# a6: pad-to-4 -- preprocess/process_dataset.py:146-150
# ---------------------------------------------------------------------------------------------
def pad_time_reflect(mel: np.ndarray, multiple: int = 4) -> np.ndarray:
    """``F.pad(mel, (0, pad), mode="reflect")`` when ``T % multiple != 0``: ``out[T + j] = mel[T - 2 - j]``."""
    t = mel.shape[-1]
    if t % multiple == 0:
        return mel
    pad = multiple - t % multiple
    if t <= pad:
        raise RuntimeError("reflect padding needs more frames than the pad")
    return np.concatenate([mel, mel[..., -2:-pad - 2:-1]], axis=-1)


# ---------------------------------------------------------------------------------------------
# a10 / a11: normalisations
# ---------------------------------------------------------------------------------------------
def normalise_global(mel: np.ndarray, mean: float = MEL_MEAN_DEFAULT, std: float = MEL_STD_DEFAULT) -> np.ndarray:
    """``(mel - mel_mean) / mel_std`` (models/modeling_vae.py:317-319)."""
    return (mel - mel.dtype.type(mean)) / mel.dtype.type(std)


def normalise_per_utterance(mel: np.ndarray, min_std: float = 1e-5) -> np.ndarray:
    """Per-bin, per-utterance ``(mel - mean_t) / clamp(std_t, 1e-5)`` with the *unbiased* std
    (eval/eval_vae.py:80-82)."""
    mean = mel.mean(axis=-1, keepdims=True)
    std = np.maximum(mel.std(axis=-1, keepdims=True, ddof=1), min_std)
    return (mel - mean) / std


# ---------------------------------------------------------------------------------------------
# a8 / a9: statistics pass -- preprocess/compute_mel_stats.py:19-36
# ---------------------------------------------------------------------------------------------
def stats_accumulate_scalar(mels: Iterable[np.ndarray]) -> Tuple[float, float, int]:
    """The reference loop: per file ``sum()`` and ``(mel**2).sum()`` in float32, accumulated across files
    as Python floats; ``numel`` as an exact integer (compute_mel_stats.py:23-28)."""
    total_sum = 0.0
    total_sq = 0.0
    total_count = 0
    for mel in mels:
        m = np.asarray(mel, dtype=np.float32)
        total_sum += float(np.sum(m, dtype=np.float32))
        total_sq += float(np.sum((m * m).astype(np.float32), dtype=np.float32))
        total_count += int(m.size)
    return total_sum, total_sq, total_count


def stats_finalise(total_sum: float, total_sq: float, count: int, var_floor: float = 1e-8) -> Tuple[float, float]:
    """``mean = S/N; var = max(S2/N - mean^2, 1e-8); std = sqrt(var)`` (compute_mel_stats.py:30-33)."""
    mean = total_sum / count
    var = max(total_sq / count - mean * mean, var_floor)
    return mean, math.sqrt(var)


def stats_per_bin(mels: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray, int]:
    """Per-bin restatement of the same moments in fp64: ``sum[b], sumsq[b]`` over all frames of all clips
    and the frame count.  The scalar statistics follow from these:
    ``S = sum_b sum[b]``, ``N = n_mels * frames``."""
    n_mels = mels[0].shape[0]
    s = np.zeros(n_mels, dtype=np.float64)
    s2 = np.zeros(n_mels, dtype=np.float64)
    frames = 0
    for mel in mels:
        m = np.asarray(mel, dtype=np.float64)
        s += m.sum(axis=1)
        s2 += (m * m).sum(axis=1)
        frames += m.shape[1]
    return s, s2, frames


def stats_per_bin_finalise(s: np.ndarray, s2: np.ndarray, frames: int, var_floor: float = 1e-8):
    mean = s / frames
    var = np.maximum(s2 / frames - mean * mean, var_floor)
    return mean, np.sqrt(var)


def format_stats_lines(mean: float, std: float) -> Tuple[str, str]:
    """The two lines the reference prints (compute_mel_stats.py:35-36; note the two spaces)."""
    return f"Global mel_mean: {mean:.6f}", f"Global mel_std:  {std:.6f}"


# ---------------------------------------------------------------------------------------------
# whole-file pipeline of process_dataset.py --mel_only for one clip (lines 140-156)
# ---------------------------------------------------------------------------------------------
def dataset_mel(wav_cl: np.ndarray, window=None, fb=None, dtype=np.float64) -> np.ndarray:
    """``[C, L]`` -> saved ``{"mel": [80, T4]}`` payload value: process_audio_chunk -> MelExtractor ->
    reflect pad of the time axis to a multiple of 4."""
    w = process_audio_chunk(wav_cl)[0]
    return pad_time_reflect(logmel(w, window, fb, dtype=dtype), 4)
