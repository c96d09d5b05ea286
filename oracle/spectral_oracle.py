"""CPU restatement (numpy) of the reference's VAE-side spectral op -- TEST INFRASTRUCTURE ONLY, never imported by the product.

``stft_mag`` follows ``AcousticVAE._stft_mag`` (/root/reference/models/modeling_vae.py:271-289): rows of ``x[B, C, T]`` are cut
into frames ``x[t * hop : t * hop + n_fft]`` (``center=False``), multiplied by the periodic Hann window, transformed with a
one-sided FFT (no normalisation) and reduced to magnitudes.  ``stft_loss`` follows ``:291-305``.  Pinned against outputs of the
unmodified reference function (tests/golden/stft_mag_cases.npz, minted by oracle/gen_golden_spectral.py).
"""
from __future__ import annotations

import numpy as np

STFT_LOSS_SPECS = ((256, 64), (128, 32), (64, 16))


def hann_periodic(n: int, dtype=np.float64) -> np.ndarray:
    """torch.hann_window(n) (periodic=True): 0.5 - 0.5 cos(2 pi k / n)."""
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)).astype(dtype)


def stft_mag(x: np.ndarray, n_fft: int, hop: int, window=None, dtype=np.float64) -> np.ndarray:
    """``x[B, C, T]`` -> ``[B, C, n_fft // 2 + 1, 1 + (T - n_fft) // hop]`` (modeling_vae.py:272-289)."""
    B, C, T = x.shape
    if T < n_fft:
        raise RuntimeError(f"expected 0 < n_fft <= {T}, but got n_fft={n_fft}")
    w = hann_periodic(n_fft) if window is None else np.asarray(window, dtype=np.float64)
    frames = 1 + (T - n_fft) // hop
    rows = x.reshape(B * C, T).astype(dtype)
    idx = np.arange(n_fft)[None, :] + hop * np.arange(frames)[:, None]           # [frames, n_fft]
    seg = rows[:, idx] * w.astype(dtype)[None, None, :]                           # [rows, frames, n_fft]
    spec = np.fft.rfft(seg, axis=-1)                                              # [rows, frames, n_freq]
    return np.abs(spec).transpose(0, 2, 1).reshape(B, C, n_fft // 2 + 1, frames)


def stft_loss(x: np.ndarray, y: np.ndarray) -> float:
    """modeling_vae.py:291-305: mean over the fitting resolutions of mean |mag(x) - mag(y)|."""
    T = x.shape[-1]
    specs = [(n, h) for n, h in STFT_LOSS_SPECS if n <= T]
    if not specs:
        return 0.0
    return float(sum(np.mean(np.abs(stft_mag(x, n, h) - stft_mag(y, n, h))) for n, h in specs) / len(specs))
