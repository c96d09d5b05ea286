"""BASELINE.json configs 1, 3, 4 and 5 (config 2 is bench.py's headline).  One JSON line per measurement.

    python tools/bench_configs.py --config 1            # single 10 s clip, batch 1: GPU latency next to the CPU reference
    python tools/bench_configs.py --config 3 [--clips N]  # stats pass over N variable-length clips (1-30 s), one all-reduce (torchrun for N GPUs)
    python tools/bench_configs.py --config 4            # ragged 0.5-20 s training-feed batch -> padded bf16 normalised log-mel + lens
    python tools/bench_configs.py --config 5            # roofline sweep: clip length x batch size

Parity at these sizes is in tests/test_gpu_parity.py; this tool measures.  Inputs are synthetic (bench.synth_batch).  Timing: CUDA events
on the launching stream, >= 3 warm-up launches, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import audio_calm_b200 as acb
from bench import SAMPLE_RATE, measured_peak, stats_pass, synth_batch


def timed(fn, steps, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        if flush is not None:
            flush.zero_()                      # > L2: evicts the previous iteration's data
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in ev]))


def emit(**kw):
    print(json.dumps(kw), flush=True)


def config1(args, fe, device):
    """Single 10 s clip, batch 1 (the reference's own call shape, process_dataset.py:109-144)."""
    from audio_calm_b200.preprocess.core import MelExtractor
    from oracle.ref_torch_port import RefMelExtractor
    L = 10 * SAMPLE_RATE
    x = synth_batch(1, L, device)
    ext = MelExtractor().to(device).eval()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    with torch.inference_mode():
        ms = timed(lambda: ext(x), 50, flush=flush)
        ms_hot = timed(lambda: ext(x), 50)
    xc = x.cpu()
    ref = RefMelExtractor().eval()
    cpu = {}
    for threads in (1, os.cpu_count() or 1):
        torch.set_num_threads(threads)
        with torch.inference_mode():
            for _ in range(3):
                ref(xc)
            t0 = time.perf_counter()
            n = 0
            while time.perf_counter() - t0 < 3.0:
                ref(xc)
                n += 1
            cpu[threads] = (time.perf_counter() - t0) / n * 1e3
    d = float((ext(x).cpu() - ref(xc)).abs().max())
    emit(config=1, workload="1 clip x 10 s -> log-mel [1, 80, 626] fp32 (MelExtractor drop-in, 1 launch)", gpu_ms_cold_l2=ms, gpu_ms_warm=ms_hot,
         gpu_audio_s_per_s=10.0 / (ms * 1e-3), cpu_ms={str(k): v for k, v in cpu.items()},
         cpu_audio_s_per_s={str(k): 10.0 / (v * 1e-3) for k, v in cpu.items()}, max_abs_diff_vs_cpu_reference=d)


def config3(args, fe, device, rank, world, dist):
    """compute_mel_stats over variable-length clips: fused extraction + per-bin fp64 moments, ONE all-reduce.  The clip set is the
    globally defined one of bench.stats_clip_set (same seeds at every world size), so the statistics of N = 1 and N = 8 compare."""
    rec = stats_pass(acb, fe, device, rank, world, dist, args.clips, per_launch=args.clips_per_launch, reps=2)
    if rank == 0:
        emit(config=3, **rec)


def config4(args, fe, device, rank, world, dist):
    """Ragged training-feed batch (0.5-20 s) -> padded bf16 scalar-normalised log-mel + valid-frame counts."""
    rng = np.random.default_rng(4 + rank)
    B = args.batch
    lens = rng.integers(SAMPLE_RATE // 2, 20 * SAMPLE_RATE + 1, size=B).astype(np.int64)
    clips = [synth_batch(1, int(n), device, seed=1000 * rank + i)[0] for i, n in enumerate(lens)]
    batch = acb.pack_clips(clips, device)
    affine = (acb.MEL_MEAN_DEFAULT, acb.MEL_STD_DEFAULT)
    feats, frames = fe.forward_ragged(batch, out_dtype=torch.bfloat16, affine=affine)
    def fresh_batch_step():                                 # a training loop sees every batch once: the tile plan is built per call
        batch.plans.clear()
        fe.forward_ragged(batch, out_dtype=torch.bfloat16, affine=affine, out=feats)
    ms = timed(fresh_batch_step, 20)
    # the launch alone: captured once in a CUDA graph (plan cached with the batch) and replayed -- no Python between launches
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fe.forward_ragged(batch, out_dtype=torch.bfloat16, affine=affine, out=feats)
    ms_graph = timed(g.replay, 50)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    audio_s = float(lens.sum()) / SAMPLE_RATE
    alg = int(4 * lens.sum() + 2 * 80 * feats.shape[2] * B)
    peak, _ = measured_peak()
    if rank == 0:
        emit(config=4, workload=f"ragged batch of {B} clips/GPU (0.5-20 s) -> [{B}, 80, {feats.shape[2]}] bf16 normalised, zero tail + lens",
             n_gpus=world, ms=ms, audio_hours_per_s=audio_s * world / 3600 / (ms * 1e-3), algorithmic_bytes=alg,
             achieved_gbs=alg / (ms * 1e-3) / 1e9, roofline_frac=alg / (ms * 1e-3) / 1e9 / peak,
             kernel_only_ms=ms_graph, kernel_only_roofline_frac=alg / (ms_graph * 1e-3) / 1e9 / peak,
             valid_frames=int(frames.sum().item()), padded_frames=int(feats.shape[0] * feats.shape[2]))


def config5(args, fe, device):
    """Roofline sweep: clip length x batch, fp32 out."""
    peak, _ = measured_peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    for sec in (1, 2, 5, 10, 20, 30, 60):
        for B in (1, 4, 16, 64, 256, 1024, 4096):
            L = sec * SAMPLE_RATE
            T = 1 + L // 256
            bytes_io = B * (4 * L + 4 * 80 * T)
            if bytes_io > 24e9:
                continue
            x = torch.randn(B, L, device=device) * 0.1
            out = torch.empty((B, 80, T), device=device)
            small = bytes_io < 200e6
            ms = timed(lambda: fe.forward(x, out=out), 5 if bytes_io > 2e9 else 20, flush=flush if small else None)
            emit(config=5, clip_seconds=sec, batch=B, ms=ms, audio_s_per_s=B * sec / (ms * 1e-3), achieved_gbs=bytes_io / (ms * 1e-3) / 1e9,
                 roofline_frac=bytes_io / (ms * 1e-3) / 1e9 / peak, l2=("flushed between launches" if small else "working set > L2"))
            del x, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True, choices=[1, 3, 4, 5])
    ap.add_argument("--clips", type=int, default=100_000)
    ap.add_argument("--clips-per-launch", type=int, default=512)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    fe = acb.LogMelFrontend(device)
    if args.config == 1 and rank == 0:
        config1(args, fe, device)
    elif args.config == 3:
        config3(args, fe, device, rank, world, dist)
    elif args.config == 4:
        config4(args, fe, device, rank, world, dist)
    elif args.config == 5 and rank == 0:
        config5(args, fe, device)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
