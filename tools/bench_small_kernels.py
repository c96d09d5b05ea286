"""Achieved HBM bandwidth of the small (memory-bound) kernels around the fused log-mel kernel.  One JSON line each.

    python tools/bench_small_kernels.py

peak_abs (per-clip max|x|, 4 B/sample read), process_audio_chunk (mixdown + peak + scale), moments over stored features (320 B/frame read),
per-utterance normalisation (2 reads + 1 write), crop_collate, pad_collate.  Working sets exceed the 126 MB L2.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import audio_calm_b200 as acb
from audio_calm_b200.preprocess.core import process_audio_chunk
from bench import measured_peak


def timed(fn, steps=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    peak, src = measured_peak()
    fe = acb.LogMelFrontend("cuda")
    out = []

    def emit(kernel, ms, nbytes, note):
        line = {"kernel": kernel, "ms": ms, "algorithmic_bytes": int(nbytes), "achieved_gbs": nbytes / (ms * 1e-3) / 1e9,
                "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peak, "peak_gbs": peak, "peak_source": src, "workload": note}
        print(json.dumps(line), flush=True)

    x = torch.randn(256, 480000, device="cuda") * 0.1
    emit("peak_abs_kernel", timed(lambda: fe.peak_abs(x)), x.numel() * 4, "256 x 30 s clips, 4 B/sample read")
    st = torch.randn(2, 160000 * 60, device="cuda") * 0.1
    emit("mixdown_peak_kernel + peak_scale_kernel", timed(lambda: process_audio_chunk(st)), st.numel() * 4 + 3 * st.shape[1] * 4,
         "one stereo clip of 600 s: 2 reads + write, then read + write")
    feats = torch.randn(256, 80, 1876, device="cuda")
    acc = acb.MelStatsAccumulator(80, "cuda")
    emit("moments_rows_kernel", timed(lambda: acc.update(feats)), feats.numel() * 4, "256 x [80, 1876] fp32 features, 320 B/frame read")
    import ctypes
    lib = acb._lib.load()
    o = torch.empty_like(feats)
    s = torch.cuda.current_stream().cuda_stream
    emit("normalize_rows_kernel", timed(lambda: lib.acb_normalize_per_utterance(feats.data_ptr(), o.data_ptr(), 256, 80, 1876, None, 1e-5, s)),
         feats.numel() * 4 * 3, "256 x [80, 1876] fp32: mean pass + variance pass + write")
    frames = torch.full((256,), 1876, dtype=torch.int64)
    emit("crop_pad_kernel", timed(lambda: acb.crop_collate(feats, frames, 256, is_eval=True)), 256 * 80 * 256 * 4 * 2,
         "256 x [80, 1876] fp32 -> [256, 80, 256] centre crops (read + write of the crop)")
    lens = torch.full((512,), 384, dtype=torch.int64)
    lat = torch.randn(512 * 384, 128, device="cuda")
    emit("pad_transpose_kernel", timed(lambda: acb.pad_collate_packed(lat, lens)), lat.numel() * 4 * 2,
         "512 x (384, 128) fp32 latents -> [512, 128, 384] (read + write)")


if __name__ == "__main__":
    main()
