"""Dataset-wide mel statistics: per-bin moments on the device, one all-reduce, finalise like the reference.

The reference's pass (``preprocess/compute_mel_stats.py:19-36``) is a single-process loop that keeps one
global ``sum``, ``sum of squares`` and ``count``.  Here every GPU keeps ``sum[b]`` and ``sumsq[b]`` per mel
bin in fp64 plus an exact integer frame count for its shard of utterances; the shards are combined with ONE
``all_reduce(SUM)`` of ``2 * n_mels + 1`` fp64 values and finalised identically on every rank.  The
reference's two scalars follow exactly from the per-bin moments
(``S = sum_b sum[b]``, ``N = n_mels * frames``).
"""
from __future__ import annotations

import ctypes
import math
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

VAR_FLOOR = 1e-8  # compute_mel_stats.py:32


@dataclass
class MelStats:
    bin_mean: np.ndarray     # [n_mels] fp64
    bin_std: np.ndarray      # [n_mels] fp64
    mel_mean: float          # the reference's "Global mel_mean"
    mel_std: float           # the reference's "Global mel_std"
    frames: int
    count: int               # n_mels * frames == sum of mel.numel() (compute_mel_stats.py:28)

    def lines(self):
        """Exactly what the reference prints (compute_mel_stats.py:35-36; two spaces after ``mel_std:``)."""
        return [f"Global mel_mean: {self.mel_mean:.6f}", f"Global mel_std:  {self.mel_std:.6f}"]

    def state(self) -> Dict[str, object]:
        """Stats-file payload: the only on-disk stats layout of the reference is ``{"mean": [D], "std": [D]}``
        (preprocess/compute_latent_stats.py:44-47); the scalar ``mel_mean`` / ``mel_std`` (pasted by hand into
        config/calm_config.yaml:62-63 in the reference) ride along."""
        return {"mean": torch.from_numpy(self.bin_mean.astype(np.float32)),
                "std": torch.from_numpy(self.bin_std.astype(np.float32)),
                "mel_mean": self.mel_mean, "mel_std": self.mel_std, "frames": self.frames, "count": self.count}

    def save(self, path: str) -> None:
        torch.save(self.state(), path)

    @classmethod
    def load(cls, path: str) -> "MelStats":
        """Read a stats file back: either one written by :meth:`save`, or the reference's bare
        ``{"mean": [D], "std": [D]}`` layout (preprocess/compute_latent_stats.py:44-47), for which the scalar
        pair is the mean of the per-bin means and the pooled standard deviation."""
        payload = torch.load(path, map_location="cpu", weights_only=False)
        if not isinstance(payload, dict) or "mean" not in payload or "std" not in payload:
            raise ValueError(f"{path}: not a stats file (expected a dict with 'mean' and 'std')")
        bm = np.asarray(torch.as_tensor(payload["mean"]).detach().to(torch.float64).reshape(-1).numpy())
        bs = np.asarray(torch.as_tensor(payload["std"]).detach().to(torch.float64).reshape(-1).numpy())
        if bm.shape != bs.shape or bm.size == 0:
            raise ValueError(f"{path}: 'mean' and 'std' must be vectors of the same length")
        if "mel_mean" in payload and "mel_std" in payload:
            mean, std = float(payload["mel_mean"]), float(payload["mel_std"])
        else:
            mean = float(bm.mean())
            std = math.sqrt(max(float((bs * bs + bm * bm).mean()) - mean * mean, VAR_FLOOR))
        frames = int(payload.get("frames", 0))
        return cls(bm, bs, mean, std, frames, int(payload.get("count", frames * bm.size)))

    def affine(self, per_bin: bool = True):
        """The ``affine=`` argument of :meth:`LogMelFrontend.forward`: per-bin ``(mean[n_mels], std[n_mels])`` fp32 tensors
        (north-star wording: "per-bin normalisation with the stored mel stats") or the reference's scalar pair
        (models/modeling_vae.py:317-319)."""
        if not per_bin:
            return (float(self.mel_mean), float(self.mel_std))
        return (torch.from_numpy(self.bin_mean.astype(np.float32)), torch.from_numpy(self.bin_std.astype(np.float32)))


def finalize_moments(moments: np.ndarray, frames: int, var_floor: float = VAR_FLOOR) -> MelStats:
    """Host-side finalise (compute_mel_stats.py:30-33 per bin and globally).  Pure numpy: also used on CPU by the
    multi-process tests; the C-ABI twin is ``acb_moments_finalize``."""
    m = np.asarray(moments, dtype=np.float64)
    n_mels = m.shape[0] // 2
    if frames <= 0:
        raise ValueError("no frames accumulated")
    s, s2 = m[:n_mels], m[n_mels:]
    bin_mean = s / frames
    bin_var = np.maximum(s2 / frames - bin_mean * bin_mean, var_floor)
    count = n_mels * frames
    mean = float(s.sum() / count)
    var = max(float(s2.sum() / count) - mean * mean, var_floor)
    return MelStats(bin_mean, np.sqrt(bin_var), mean, math.sqrt(var), int(frames), int(count))


class MelStatsAccumulator:
    """Running per-bin moments on one device (fp64) + exact frame count on the host."""

    def __init__(self, n_mels: int = 80, device="cuda"):
        self.n_mels = int(n_mels)
        self.device = torch.device(device)
        self.moments = torch.zeros(2 * self.n_mels, dtype=torch.float64, device=self.device)
        self.frames = 0
        self._ws: Optional[torch.Tensor] = None   # scratch of acb_moments_accumulate, allocated on first use

    # -- S1: moments of features that already exist (files written by the dataset driver) --
    def update(self, feat: torch.Tensor, frames: Optional[torch.Tensor] = None) -> None:
        """``feat[B, n_mels, cap]`` (or ``[n_mels, T]``) fp32/bf16 on the device; ``frames[B]`` int64 valid frames."""
        from . import _lib
        lib = _lib.load()
        if feat.dim() == 2:
            feat = feat.unsqueeze(0)
        if not feat.is_cuda:
            raise RuntimeError("MelStatsAccumulator.update expects device features (no CPU fallback)")
        if feat.dtype not in (torch.float32, torch.bfloat16):
            feat = feat.float()
        feat = feat.contiguous()
        B, M, cap = (int(x) for x in feat.shape)
        if M != self.n_mels:
            raise ValueError(f"expected {self.n_mels} mel bins, got {M}")
        fr_ptr = None
        if frames is not None:
            frames = frames.to(feat.device, torch.int64).contiguous()
            fr_ptr = frames.data_ptr()
        dtype = _lib.ACB_F32 if feat.dtype == torch.float32 else _lib.ACB_BF16
        if self._ws is None:
            self._ws = torch.empty(int(lib.acb_moments_accumulate_workspace_bytes(self.n_mels)) // 8, dtype=torch.float64, device=self.device)
        _lib.check(lib.acb_moments_accumulate(feat.data_ptr(), dtype, B, M, cap, M * cap, fr_ptr, self.moments.data_ptr(),
                                              self._ws.data_ptr(), torch.cuda.current_stream(feat.device).cuda_stream),
                   "acb_moments_accumulate")
        self.frames += int(frames.sum().item()) if frames is not None else B * cap

    # -- combine shards: one collective --
    def all_reduce(self, group=None) -> None:
        """SUM the per-rank moments and frame counts across the process group with a single ``all_reduce`` of
        ``2 * n_mels + 1`` fp64 values (the count is exact in fp64 below 2^53 frames)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        buf = torch.cat([self.moments, torch.tensor([float(self.frames)], dtype=torch.float64, device=self.moments.device)])
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        self.moments = buf[:-1].clone()
        self.frames = int(round(buf[-1].item()))

    def finalize(self, var_floor: float = VAR_FLOOR) -> MelStats:
        return finalize_moments(self.moments.detach().cpu().numpy(), self.frames, var_floor)


def normalize_per_utterance(mel: torch.Tensor, frames: Optional[torch.Tensor] = None, min_std: float = 1e-5,
                            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``(mel - mean_t) / clamp(std_t, min=1e-5)`` per utterance and per mel bin, unbiased std over time -- the three
    lines of the reference's eval path (eval/eval_vae.py:80-82) as one kernel (``acb_normalize_per_utterance``).

    ``mel``: device fp32 ``[..., n_mels, T]``; ``frames[B]`` (int64) limits a padded batch to each clip's valid frames."""
    from . import _lib
    lib = _lib.load()
    if not mel.is_cuda:
        raise RuntimeError("normalize_per_utterance expects a CUDA tensor (no CPU fallback)")
    if mel.dtype != torch.float32 or mel.dim() < 2:
        raise ValueError("normalize_per_utterance expects float32 [..., n_mels, T]")
    x = mel.contiguous()
    n_mels, cap = int(x.shape[-2]), int(x.shape[-1])
    B = x.numel() // max(n_mels * cap, 1)
    if out is None:
        out = torch.empty_like(x)
    elif out.shape != x.shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != x.device:
        raise ValueError("out must be a contiguous float32 tensor of the input's shape on the same device")
    fr_ptr = None
    if frames is not None:
        frames = frames.to(x.device, torch.int64).contiguous()
        if frames.numel() != B:
            raise ValueError("frames must hold one count per utterance")
        fr_ptr = frames.data_ptr()
    if B and cap:
        with torch.cuda.device(x.device):
            _lib.check(lib.acb_normalize_per_utterance(x.data_ptr(), out.data_ptr(), B, n_mels, cap, fr_ptr, float(min_std),
                                                       torch.cuda.current_stream(x.device).cuda_stream), "acb_normalize_per_utterance")
    return out


def finalize_moments_c(moments: np.ndarray, frames: int, var_floor: float = VAR_FLOOR) -> MelStats:
    """Same as :func:`finalize_moments` through ``acb_moments_finalize`` (host function of the C ABI)."""
    from . import _lib
    lib = _lib.load()
    m = np.ascontiguousarray(moments, dtype=np.float64)
    n_mels = m.shape[0] // 2
    bm = np.zeros(n_mels)
    bs = np.zeros(n_mels)
    gm, gs = ctypes.c_double(), ctypes.c_double()
    _lib.check(lib.acb_moments_finalize(m.ctypes.data, n_mels, int(frames), float(var_floor), bm.ctypes.data, bs.ctypes.data,
                                        ctypes.cast(ctypes.byref(gm), ctypes.c_void_p), ctypes.cast(ctypes.byref(gs), ctypes.c_void_p)),
               "acb_moments_finalize")
    return MelStats(bm, bs, gm.value, gs.value, int(frames), int(n_mels * frames))
