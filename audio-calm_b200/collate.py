"""Training-feed collation of stored features on the device.

The reference builds its batches on the host, one item at a time:

* VAE training (``train/train_vae.py:83-116``): ``MelDataset.__getitem__`` crops a ``[80, T]`` mel to ``crop_size = 256`` frames
  (random start for training -- ``torch.randint(0, T - crop, (1,))`` -- or centred for eval) or zero-pads it on the right, and
  ``data_collator`` stacks the items into ``{"mel": [B, 80, 256], "labels": same tensor}``.
* CALM training (``train/train_calm.py:178-221``): ``CalmCollator`` pads ragged ``(T_i, D)`` latents with
  ``pad_sequence(batch_first=True, padding_value=audio_pad_val)`` and transposes to channels-first ``(B, D, T_max)``, next to
  ``audio_lens``; in ASR training a random span of 5-10 frames is zeroed first (``_apply_spec_augment``, ``:184-191``).

Here both are one launch over the whole batch (``acb_crop_pad`` / ``acb_pad_transpose``).  Random draws stay on the host side of the
call (start / mask tensors), so the kernels are deterministic and testable.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib

CROP_SIZE_DEFAULT = 256   # config/vae_config.yaml crop_size; MelDataset(crop_size=...)


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _lib.ACB_F32
    if t.dtype == torch.bfloat16:
        return _lib.ACB_BF16
    raise RuntimeError(f"expected float32 or bfloat16 features, got {t.dtype}")


def crop_starts(frames: torch.Tensor, crop_size: int, is_eval: bool, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Start frame of every clip's crop, with the reference's rules (train_vae.py:86-102): clips longer than ``crop_size`` start at
    ``randint(0, T - crop)`` (upper bound exclusive, as ``torch.randint``) in training and at ``(T - crop) // 2`` in eval; shorter
    clips start at 0 and are zero-padded.  ``frames`` may live on the host or the device; the result follows it."""
    room = (frames.to(torch.int64) - int(crop_size)).clamp_(min=0)
    if is_eval:
        return room // 2
    u = torch.rand(tuple(frames.shape), generator=generator, dtype=torch.float64).to(frames.device)   # host draw (CPU generator)
    return torch.minimum((u * room.to(torch.float64)).to(torch.int64), (room - 1).clamp_(min=0))


def crop_collate(feat: torch.Tensor, frames: Optional[torch.Tensor] = None, crop_size: int = CROP_SIZE_DEFAULT, is_eval: bool = False,
                 start: Optional[torch.Tensor] = None, pad_value: float = 0.0,
                 generator: Optional[torch.Generator] = None) -> Dict[str, torch.Tensor]:
    """``feat[B, n_mels, cap]`` (device, fp32 / bf16; e.g. the output of ``LogMelFrontend.forward_ragged``) with ``frames[B]`` valid
    frames -> ``{"mel": [B, n_mels, crop_size], "labels": same tensor, "start": [B]}`` like ``MelDataset`` + ``data_collator``."""
    if not feat.is_cuda:
        raise RuntimeError("crop_collate expects device features (no CPU fallback)")
    if feat.dim() != 3 or feat.stride(2) != 1 or feat.stride(1) != feat.shape[2]:
        feat = feat.contiguous()
    B, M, cap = (int(x) for x in feat.shape)
    dev = feat.device
    if frames is None:
        frames = torch.full((B,), cap, dtype=torch.int64, device=dev)
    frames = frames.to(dev, torch.int64).contiguous()
    if start is None:
        start = crop_starts(frames, crop_size, is_eval, generator)
    start = start.to(dev, torch.int64).contiguous()
    out = torch.empty((B, M, int(crop_size)), dtype=feat.dtype, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.acb_crop_pad(feat.data_ptr(), _dt(feat), B, M, cap, int(feat.stride(0)), frames.data_ptr(), start.data_ptr(),
                                    out.data_ptr(), int(crop_size), float(pad_value), torch.cuda.current_stream(dev).cuda_stream),
                   "acb_crop_pad")
    return {"mel": out, "labels": out, "start": start}


def pad_collate(items: Sequence[torch.Tensor], pad_value: float = 0.0, mask: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                out_frames: Optional[int] = None, device: Optional[torch.device] = None) -> Dict[str, torch.Tensor]:
    """Ragged ``(T_i, D)`` items (host or device, fp32 / bf16) -> ``{"audio_features": [B, D, T_max], "audio_lens": [B] int64}`` as
    ``CalmCollator`` builds them (train_calm.py:205-215).  ``mask = (start[B], length[B])`` zeroes that span of frames per clip
    (the SpecAugment of ``:184-191``; pass length 0 for clips that are not masked)."""
    if not items:
        raise ValueError("pad_collate needs at least one item")
    D = int(items[0].shape[1])
    dtype = items[0].dtype
    dev = torch.device(device) if device is not None else (items[0].device if items[0].is_cuda else torch.device("cuda", torch.cuda.current_device()))
    lens = [int(x.shape[0]) for x in items]
    flat = torch.cat([x.reshape(-1, D).to(dev, non_blocking=True) for x in items], dim=0).contiguous()
    return pad_collate_packed(flat, torch.tensor(lens, dtype=torch.int64), pad_value, mask, out_frames)


def pad_collate_packed(flat: torch.Tensor, lens: torch.Tensor, pad_value: float = 0.0,
                       mask: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, out_frames: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Same as :func:`pad_collate` for features already packed on the device as time-major rows ``flat[sum T_i, D]``."""
    if not flat.is_cuda:
        raise RuntimeError("pad_collate_packed expects device features (no CPU fallback)")
    dev = flat.device
    D = int(flat.shape[1])
    lens_host = lens.to("cpu", torch.int64)
    B = int(lens_host.shape[0])
    if int(lens_host.sum()) != int(flat.shape[0]):
        raise ValueError("lens do not add up to the number of rows")
    T = int(out_frames) if out_frames is not None else int(lens_host.max())
    offs = torch.zeros(B, dtype=torch.int64)
    offs[1:] = torch.cumsum(lens_host, 0)[:-1]
    lens_d, offs_d = lens_host.to(dev, non_blocking=True), offs.to(dev, non_blocking=True)
    m0 = m1 = None
    if mask is not None:
        m0 = mask[0].to(dev, torch.int64).contiguous()
        m1 = mask[1].to(dev, torch.int64).contiguous()
    out = torch.empty((B, D, T), dtype=flat.dtype, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.acb_pad_transpose(flat.data_ptr(), _dt(flat), offs_d.data_ptr(), lens_d.data_ptr(), B, D, out.data_ptr(), T,
                                         float(pad_value), None if m0 is None else m0.data_ptr(), None if m1 is None else m1.data_ptr(),
                                         torch.cuda.current_stream(dev).cuda_stream), "acb_pad_transpose")
    return {"audio_features": out, "audio_lens": lens_host.clone()}
