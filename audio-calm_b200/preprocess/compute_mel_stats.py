"""Drop-in for the reference's ``preprocess/compute_mel_stats.py`` (same CLI, same two printed lines).

Reference algorithm (``compute_mel_stats.py:19-36``): walk ``--root`` for ``*.pt``, and over every ``payload["mel"]``
accumulate sum, sum of squares and element count; print ``Global mel_mean`` / ``Global mel_std``.  Here the files are
loaded on the host (disk I/O is outside the hot path), shipped to the GPU in padded batches and reduced by
``acb_moments_accumulate`` into per-bin fp64 moments; under ``torchrun`` every rank takes a shard of the files and the
moments are combined with ONE all-reduce.  The scalars the reference prints follow exactly from the per-bin moments.

    python preprocess/compute_mel_stats.py --root data/mels/train [--save mel_stats.pt] [--per_bin]
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

try:
    from ..stats import MelStatsAccumulator
    from ..sharding import contiguous_shard
except ImportError:  # run as a script / top-level module with the package directory on sys.path
    _root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if _root not in sys.path:
        sys.path.insert(0, _root)
    from audio_calm_b200.stats import MelStatsAccumulator
    from audio_calm_b200.sharding import contiguous_shard


def iter_mel_files(root):
    """Same generator as the reference (compute_mel_stats.py:7-11): every ``*.pt`` under ``root``."""
    for dirpath, _, filenames in os.walk(root):
        for name in filenames:
            if name.endswith(".pt"):
                yield os.path.join(dirpath, name)


def _flush(acc, mels):
    if not mels:
        return
    n_mels = mels[0].shape[0]
    cap = max(m.shape[1] for m in mels)
    host = torch.zeros((len(mels), n_mels, cap), dtype=torch.float32, pin_memory=True)
    for i, m in enumerate(mels):
        host[i, :, :m.shape[1]] = m
    frames = torch.tensor([m.shape[1] for m in mels], dtype=torch.int64)
    acc.update(host.to(acc.device, non_blocking=True), frames)
    mels.clear()


def compute_stats(files, device="cuda", batch_frames=200_000, progress=False):
    """Per-bin moments of ``payload["mel"]`` over ``files`` -> MelStatsAccumulator (not yet all-reduced)."""
    acc = None
    pending, pending_frames = [], 0
    it = files
    if progress:
        try:
            from tqdm import tqdm
            it = tqdm(files, desc="Scanning mels")
        except ImportError:
            pass
    for path in it:
        payload = torch.load(path, map_location="cpu", weights_only=False)
        mel = payload["mel"].float()                                     # [80, T]  (compute_mel_stats.py:25)
        if acc is None:
            acc = MelStatsAccumulator(int(mel.shape[0]), device)
        pending.append(mel)
        pending_frames += int(mel.shape[1])
        if pending_frames >= batch_frames:
            _flush(acc, pending)
            pending_frames = 0
    if acc is None:
        acc = MelStatsAccumulator(80, device)
    _flush(acc, pending)
    return acc


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--root", type=str, required=True, help="root directory written by process_dataset.py --mel_only")
    parser.add_argument("--save", type=str, default=None, help="optional stats file: {'mean':[80],'std':[80],'mel_mean','mel_std'}")
    parser.add_argument("--per_bin", action="store_true", help="also print the per-bin mean/std")
    args = parser.parse_args(argv)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("compute_mel_stats (B200 build) needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    files = sorted(iter_mel_files(args.root))
    mine = [files[i] for i in contiguous_shard(len(files), rank, world)]
    acc = compute_stats(mine, torch.device("cuda", local_rank), progress=(rank == 0))
    acc.all_reduce()
    if acc.frames == 0:
        raise SystemExit(f"no .pt files with frames under {args.root}")
    st = acc.finalize()
    if rank == 0:
        for line in st.lines():                                           # compute_mel_stats.py:35-36
            print(line)
        if args.per_bin:
            for b in range(len(st.bin_mean)):
                print(f"bin {b:3d}: mean {st.bin_mean[b]:.6f} std {st.bin_std[b]:.6f}")
        if args.save:
            st.save(args.save)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return st


if __name__ == "__main__":
    main()
