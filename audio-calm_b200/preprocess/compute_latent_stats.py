"""Latent statistics pass (reference ``preprocess/compute_latent_stats.py:9-47``) on the same moments kernel.

The reference script is a hard-coded loop over ``{"latent"}`` files: it transposes ``(D, T)`` payloads whose first
dimension is one of 64/80/128/192 to ``(T, D)``, then accumulates ``sum`` / ``sum of squares`` either over everything
(``reduce_dim=True`` -> two scalars printed as ``latent_mean`` / ``latent_std``) or per dimension
(``reduce_dim=False`` -> ``latent_stats.pt = {"mean": [D], "std": [D]}``, the only stats file the reference writes).
Differences kept: variance floor 1e-12 (``:40``) and the print formats.
"""
from __future__ import annotations

import argparse
import math
import os
import sys
from glob import glob

import numpy as np
import torch

try:
    from ..stats import MelStatsAccumulator
except ImportError:
    _root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if _root not in sys.path:
        sys.path.insert(0, _root)
    from audio_calm_b200.stats import MelStatsAccumulator

VAR_FLOOR = 1e-12  # compute_latent_stats.py:40


def load_latent_dt(path):
    """Payload -> ``(D, T)`` float tensor (the reference normalises to ``(T, D)``; ``:21-25``)."""
    payload = torch.load(path, map_location="cpu", weights_only=True)
    lat = payload.get("latent", payload) if isinstance(payload, dict) else payload
    if lat.dim() == 2 and lat.shape[0] in (64, 80, 128, 192):
        return lat.float()                       # already (D, T)
    return lat.float().transpose(0, 1)           # (T, D) -> (D, T)


def latent_stats(files, device="cuda", reduce_dim=True, batch_frames=400_000):
    acc, pending, frames = None, [], 0

    def flush():
        nonlocal frames
        if not pending:
            return
        D = pending[0].shape[0]
        cap = max(x.shape[1] for x in pending)
        host = torch.zeros((len(pending), D, cap), dtype=torch.float32)
        for i, x in enumerate(pending):
            host[i, :, :x.shape[1]] = x
        acc.update(host.to(acc.device), torch.tensor([x.shape[1] for x in pending], dtype=torch.int64))
        pending.clear()
        frames = 0

    for f in files:
        lat = load_latent_dt(f)
        if acc is None:
            acc = MelStatsAccumulator(int(lat.shape[0]), device)
        pending.append(lat)
        frames += int(lat.shape[1])
        if frames >= batch_frames:
            flush()
    if acc is None:
        raise AssertionError("No .pt found")      # compute_latent_stats.py:13
    flush()
    m = acc.moments.cpu().numpy()
    D = acc.n_mels
    if reduce_dim:
        count = acc.frames * D
        mean = m[:D].sum() / count
        var = max(m[D:].sum() / count - mean * mean, VAR_FLOOR)
        return float(mean), math.sqrt(var)
    mean = m[:D] / acc.frames
    var = np.maximum(m[D:] / acc.frames - mean * mean, VAR_FLOOR)
    return torch.from_numpy(mean.astype(np.float32)), torch.from_numpy(np.sqrt(var).astype(np.float32))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--latent_dir", required=True)
    ap.add_argument("--max_files", type=int, default=None)
    ap.add_argument("--per_dim", action="store_true", help="reduce_dim=False: per-dimension stats saved to latent_stats.pt")
    ap.add_argument("--out", default="latent_stats.pt")
    args = ap.parse_args(argv)
    files = sorted(glob(os.path.join(args.latent_dir, "**", "*.pt"), recursive=True))
    if args.max_files:
        files = files[:args.max_files]
    assert files, "No .pt found"
    if not torch.cuda.is_available():
        raise RuntimeError("compute_latent_stats (B200 build) needs a CUDA device: there is no CPU fallback")
    mean, std = latent_stats(files, "cuda", reduce_dim=not args.per_dim)
    if not args.per_dim:
        print(f"latent_mean: {mean:.6f}")
        print(f"latent_std : {std:.6f}")
    else:
        print("latent_mean shape:", mean.shape)
        print("latent_std  shape:", std.shape)
        torch.save({"mean": mean, "std": std}, args.out)
        print(f"saved to {args.out}")


if __name__ == "__main__":
    main()
