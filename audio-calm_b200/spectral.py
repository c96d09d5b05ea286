"""VAE-side spectral ops on the device.

``stft_mag`` mirrors the reference's ``AcousticVAE._stft_mag`` (``models/modeling_vae.py:271-289``): short-time Fourier
magnitudes over the TIME axis of ``[B, C, T]`` features -- ``torch.stft(n_fft, hop_length, window=hann_window(win_length),
center=False, normalized=False)`` followed by ``torch.abs`` -- as one launch of ``acb_stft_mag`` instead of a framing copy, a
window multiply, a cuFFT call and an ``abs`` pass.  ``multires_stft_mags`` and ``stft_loss`` follow ``stft_loss`` (``:291-305``):
the resolutions ``(256, 64), (128, 32), (64, 16)`` that fit the sequence, L1 distance of the magnitudes, averaged.

Forward only: the loss value is what evaluation needs; training through it (autograd) is not wired yet.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import _lib

STFT_LOSS_SPECS = ((256, 64), (128, 32), (64, 16))      # models/modeling_vae.py:294
SUPPORTED_N_FFT = (64, 128, 256, 512, 1024)

_windows: Dict[Tuple[int, str], torch.Tensor] = {}


def _window(n_fft: int, device: torch.device) -> torch.Tensor:
    """``torch.hann_window(n_fft)`` (periodic), built by the same torch call as the reference's so that it is bit-identical."""
    key = (n_fft, str(device))
    w = _windows.get(key)
    if w is None:
        w = torch.hann_window(n_fft, dtype=torch.float32).to(device)
        _windows[key] = w
    return w


def stft_frames(length: int, n_fft: int, hop_length: int) -> int:
    """``1 + (T - n_fft) // hop`` frames of ``torch.stft(center=False)``."""
    if length < n_fft:
        raise RuntimeError(f"stft_mag: expected 0 < n_fft <= {length}, but got n_fft={n_fft}")      # torch.stft's complaint
    return 1 + (length - n_fft) // hop_length


def stft_mag(x: torch.Tensor, n_fft: int = 1024, hop_length: int = 256, win_length: Optional[int] = None) -> torch.Tensor:
    """``x[B, C, T]`` (device; any float dtype, computed in float32 like the reference's ``.float()``) ->
    ``[B, C, n_fft // 2 + 1, frames]`` float32 magnitudes."""
    if x.dim() != 3:
        raise ValueError("stft_mag expects [B, C, T]")
    if not x.is_cuda:
        raise RuntimeError("stft_mag (B200 build) needs a CUDA tensor: there is no CPU fallback")
    if win_length not in (None, n_fft):
        raise NotImplementedError("stft_mag: win_length other than n_fft (the reference never passes one)")
    if n_fft not in SUPPORTED_N_FFT:
        raise NotImplementedError(f"stft_mag: n_fft must be one of {SUPPORTED_N_FFT}")
    B, C, T = (int(v) for v in x.shape)
    frames = stft_frames(T, n_fft, hop_length)
    x2 = x.reshape(B * C, T).float().contiguous()
    out = torch.empty((B, C, n_fft // 2 + 1, frames), dtype=torch.float32, device=x.device)
    if B * C:
        lib = _lib.load()
        with torch.cuda.device(x.device):
            _lib.check(lib.acb_stft_mag(x2.data_ptr(), B * C, T, int(n_fft), int(hop_length), _window(n_fft, x.device).data_ptr(),
                                        out.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream), "acb_stft_mag")
    return out


def multires_stft_mags(x: torch.Tensor) -> List[torch.Tensor]:
    """Magnitudes at every resolution of ``stft_loss`` that fits ``T`` (``n_fft <= T``, models/modeling_vae.py:295)."""
    T = int(x.shape[-1])
    return [stft_mag(x, n_fft=n, hop_length=h) for n, h in STFT_LOSS_SPECS if n <= T]


def stft_loss(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """Forward value of the reference's multi-resolution STFT loss (models/modeling_vae.py:291-305): mean over the resolutions of
    the L1 distance between the magnitudes of ``x`` and ``y``; zero when no resolution fits."""
    mx, my = multires_stft_mags(x), multires_stft_mags(y)
    if not mx:
        return torch.tensor(0.0, device=x.device, dtype=x.dtype)
    loss = sum((a - b).abs().mean() for a, b in zip(mx, my))
    return loss / len(mx)
