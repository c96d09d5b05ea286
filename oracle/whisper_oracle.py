"""CPU oracle of the Whisper-style log-mel preset (the north star's wording of the front-end).

TEST INFRASTRUCTURE ONLY: imported by ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of the bench tools,
never by the product package.

Where the algorithm lives: not in the reference's own files -- the reference reaches it through the Hugging Face ASR pipeline
that scores generated speech (``eval/eval_calm.py:548-552``: ``pipeline("automatic-speech-recognition",
model="openai/whisper-tiny.en")``), i.e. the third-party, un-vendored ``transformers.WhisperFeatureExtractor``
(``transformers`` is unpinned in ``requirements/``; this container has 5.5.0), and it ships the filterbank that extractor uses
as ``models/mel_filters.npz`` (key ``mel_80``, shape (80, 201)).  The published algorithm
(``WhisperFeatureExtractor._np_extract_fbank_features``) restated here in numpy fp64:

    pad / trim the clip to 480000 samples (zeros)            -- __call__, padding="max_length"
    reflect-pad 200, frames of 400 at hop 160, periodic Hann  -- spectrogram(frame_length=400, hop_length=160, center=True)
    |rfft|^2, drop the last frame                             -- power=2.0; log_spec[:, :-1]
    mel = filters[201, 80]^T . power                          -- mel_filters (slaney scale, slaney norm, 0-8000 Hz)
    log10(max(mel, 1e-10))                                    -- log_mel="log10", mel_floor=1e-10
    max(x, x.max() - 8.0); (x + 4.0) / 4.0

Parity pinning: ``oracle/gen_golden_whisper.py`` runs the unmodified ``WhisperFeatureExtractor`` in the build container and commits
its outputs under ``tests/golden/whisper_cases.npz`` (plus ``models/mel_filters.npz``-derived checks); ``tests/test_whisper_oracle.py``
checks this module against every one of them.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

N_FFT = 400
HOP = 160
N_MELS = 80
N_SAMPLES = 480000
SAMPLE_RATE = 16000


def hann_window_f64(n: int = N_FFT) -> np.ndarray:
    """Periodic Hann (``window_function(400, "hann")`` in transformers.audio_utils; same as ``torch.hann_window(400)``)."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = 3.0 * f / 200.0
    log = 15.0 + np.log(np.maximum(f, 1e-300) / 1000.0) * (27.0 / np.log(6.4))
    return np.where(f >= 1000.0, log, lin)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    lin = 200.0 * m / 3.0
    log = 1000.0 * np.exp(np.log(6.4) / 27.0 * (m - 15.0))
    return np.where(m >= 15.0, log, lin)


def mel_filters_f64(n_mels: int = N_MELS, n_freqs: int = N_FFT // 2 + 1, sample_rate: int = SAMPLE_RATE) -> np.ndarray:
    """``[n_freqs, n_mels]`` slaney-scale, slaney-normalised triangular bank over 0 .. sr/2 (``mel_filter_bank(201, 80, 0, 8000,
    16000, norm="slaney", mel_scale="slaney")``; equals ``models/mel_filters.npz['mel_80'].T`` to 1.3e-9)."""
    fft_freqs = np.linspace(0.0, sample_rate / 2.0, n_freqs)
    mel_pts = np.linspace(_hz_to_mel(0.0), _hz_to_mel(sample_rate / 2.0), n_mels + 2)
    f_pts = _mel_to_hz(mel_pts)
    f_diff = np.diff(f_pts)
    slopes = f_pts[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    fb *= (2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels]))[None, :]
    return fb


def frames_for_length(length: int, drop_last: bool = True) -> int:
    """``1 + L // 160`` centred frames, minus the dropped last one; raises like ``np.pad(mode="reflect")`` needs L > 200."""
    if length <= N_FFT // 2:
        raise RuntimeError(f"reflect padding of {N_FFT // 2} needs more than {N_FFT // 2} samples, got {length}")
    return 1 + length // HOP - (1 if drop_last else 0)


def pad_or_trim(wav: np.ndarray, n_samples: int = N_SAMPLES) -> np.ndarray:
    wav = np.asarray(wav).reshape(-1)[:n_samples]
    out = np.zeros(n_samples, dtype=wav.dtype)
    out[: wav.shape[0]] = wav
    return out


def power_spectrogram(wav: np.ndarray, window: Optional[np.ndarray] = None) -> np.ndarray:
    """``[T, 201]`` fp64 power spectrum of every centred frame (T = 1 + L // 160)."""
    x = np.asarray(wav, dtype=np.float64).reshape(-1)
    frames_for_length(x.shape[0], drop_last=False)
    w = hann_window_f64() if window is None else np.asarray(window, dtype=np.float64)
    xp = np.pad(x, (N_FFT // 2, N_FFT // 2), mode="reflect")
    t = 1 + x.shape[0] // HOP
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(t)[:, None]
    return np.abs(np.fft.rfft(xp[idx] * w[None, :], axis=1)) ** 2


def whisper_logmel(wav: np.ndarray, window: Optional[np.ndarray] = None, fb: Optional[np.ndarray] = None, drop_last: bool = True,
                   clamp_min: float = 1e-10, dyn_range: Optional[float] = 8.0, affine: bool = True) -> np.ndarray:
    """``[80, T]`` fp64 features of one clip at its own length (no 30 s padding; see ``pad_or_trim``)."""
    p = power_spectrogram(wav, window)
    if drop_last:
        p = p[:-1]
    f = mel_filters_f64() if fb is None else np.asarray(fb, dtype=np.float64)
    x = np.log10(np.maximum(p @ f, clamp_min))
    if dyn_range is not None:
        x = np.maximum(x, x.max() - dyn_range)
    if affine:
        x = (x + 4.0) / 4.0
    return x.T.copy()
