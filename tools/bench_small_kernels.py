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
    """Device time per call: the call is captured ONCE in a CUDA graph (the C ABI launches on torch's current stream, which is
    the capture stream inside torch.cuda.graph) and replayed, so the 20-40 us of Python / ctypes work per call stay out."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g.replay()
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
    feats_big = torch.randn(2048, 80, 1876, device="cuda")
    emit("moments_rows_kernel (large)", timed(lambda: acc.update(feats_big)), feats_big.numel() * 4, "2048 x [80, 1876] fp32 features (1.2 GB read)")
    del feats_big
    o = torch.empty_like(feats)
    cs = lambda: torch.cuda.current_stream().cuda_stream      # inside a graph capture this is the capture stream
    emit("normalize_rows_kernel", timed(lambda: lib.acb_normalize_per_utterance(feats.data_ptr(), o.data_ptr(), 256, 80, 1876, None, 1e-5, cs())),
         feats.numel() * 4 * 2, "256 x [80, 1876] fp32: one read (a warp keeps its row in registers for the mean and the variance) + one write")
    # the two collation kernels are timed through the raw C ABI with every argument already on the device (the Python wrappers
    # compute starts / offsets on the host, which is not kernel time)
    frames = torch.full((256,), 1876, dtype=torch.int64, device="cuda")
    start = (frames - 256) // 2
    crop = torch.empty((256, 80, 256), device="cuda")
    emit("crop_pad_kernel", timed(lambda: lib.acb_crop_pad(feats.data_ptr(), 0, 256, 80, 1876, 80 * 1876, frames.data_ptr(), start.data_ptr(),
                                                           crop.data_ptr(), 256, 0.0, cs())), 256 * 80 * 256 * 4 * 2,
         "256 x [80, 1876] fp32 -> [256, 80, 256] centre crops (read + write of the crop; 42 MB: L2-resident, launch-bound)")
    big = torch.randn(2048, 80, 1876, device="cuda")
    frames_b = torch.full((2048,), 1876, dtype=torch.int64, device="cuda")
    start_b = (frames_b - 1024) // 2
    crop_b = torch.empty((2048, 80, 1024), device="cuda")
    emit("crop_pad_kernel (large)", timed(lambda: lib.acb_crop_pad(big.data_ptr(), 0, 2048, 80, 1876, 80 * 1876, frames_b.data_ptr(), start_b.data_ptr(),
                                                                   crop_b.data_ptr(), 1024, 0.0, cs())), 2048 * 80 * 1024 * 4 * 2,
         "2048 x [80, 1876] fp32 -> [2048, 80, 1024] centre crops (1.3 GB moved)")
    del big, crop_b
    nb = 4096
    lens_b = torch.full((nb,), 384, dtype=torch.int64, device="cuda")
    offs_b = torch.arange(nb, dtype=torch.int64, device="cuda") * 384
    lat_b = torch.randn(nb * 384, 128, device="cuda")
    pt_b = torch.empty((nb, 128, 384), device="cuda")
    emit("pad_transpose_kernel (large)", timed(lambda: lib.acb_pad_transpose(lat_b.data_ptr(), 0, offs_b.data_ptr(), lens_b.data_ptr(), nb, 128, pt_b.data_ptr(), 384,
                                                                             0.0, None, None, cs())), lat_b.numel() * 4 * 2,
         "4096 x (384, 128) fp32 latents -> [4096, 128, 384] (1.6 GB moved)")
    del lat_b, pt_b
    lens = torch.full((512,), 384, dtype=torch.int64, device="cuda")
    offs = torch.arange(512, dtype=torch.int64, device="cuda") * 384
    lat = torch.randn(512 * 384, 128, device="cuda")
    pt = torch.empty((512, 128, 384), device="cuda")
    emit("pad_transpose_kernel", timed(lambda: lib.acb_pad_transpose(lat.data_ptr(), 0, offs.data_ptr(), lens.data_ptr(), 512, 128, pt.data_ptr(), 384,
                                                                     0.0, None, None, cs())), lat.numel() * 4 * 2,
         "512 x (384, 128) fp32 latents -> [512, 128, 384] (read + write)")


    # VAE-side STFT magnitudes over the mel time axis (models/modeling_vae.py:271-305): a 4x training batch so that the working set exceeds L2
    from audio_calm_b200 import spectral
    xs = torch.randn(1024, 80, 256, device="cuda") * 3.0 - 6.0
    for n_fft, hop in spectral.STFT_LOSS_SPECS:
        frames = spectral.stft_frames(256, n_fft, hop)
        o = torch.empty((1024, 80, n_fft // 2 + 1, frames), device="cuda")
        w = spectral._window(n_fft, xs.device)
        x2 = xs.reshape(-1, 256)
        ref_ms = timed(lambda: torch.stft(x2, n_fft=n_fft, hop_length=hop, win_length=n_fft, window=w, return_complex=True, normalized=False,
                                          center=False).abs())
        emit(f"stft_mag_kernel (n_fft {n_fft}, hop {hop})",
             timed(lambda: lib.acb_stft_mag(xs.data_ptr(), 1024 * 80, 256, n_fft, hop, w.data_ptr(), o.data_ptr(), cs())),
             xs.numel() * 4 + o.numel() * 4, f"[1024, 80, 256] fp32 -> [1024, 80, {n_fft // 2 + 1}, {frames}] magnitudes (read + write); "
             f"torch.stft + abs on the same device (the reference's _stft_mag, cuFFT): {ref_ms:.4f} ms")

    # backward of the same op (the loss is trained through): ours = recomputed transform + one more FFT per frame pair + overlap-add
    # in shared memory; the reference = torch autograd through torch.stft(...).abs() (cuFFT C2R + elementwise + col2im)
    for n_fft, hop in spectral.STFT_LOSS_SPECS:
        frames = spectral.stft_frames(256, n_fft, hop)
        w = spectral._window(n_fft, xs.device)
        g = torch.randn((1024, 80, n_fft // 2 + 1, frames), device="cuda")
        gx = torch.empty_like(xs)
        x2 = xs.reshape(-1, 256).clone().requires_grad_(True)
        g2 = g.reshape(-1, n_fft // 2 + 1, frames)

        def torch_fwd_bwd():
            x2.grad = None
            torch.stft(x2, n_fft=n_fft, hop_length=hop, win_length=n_fft, window=w, return_complex=True, normalized=False, center=False).abs().backward(g2)

        for _ in range(3):
            torch_fwd_bwd()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            torch_fwd_bwd()
        e1.record()
        torch.cuda.synchronize()
        ref_ms = e0.elapsed_time(e1) / 10
        emit(f"stft_mag_backward_kernel (n_fft {n_fft}, hop {hop})",
             timed(lambda: lib.acb_stft_mag_backward(xs.data_ptr(), g.data_ptr(), 1024 * 80, 256, n_fft, hop, w.data_ptr(), gx.data_ptr(), cs())),
             xs.numel() * 4 * 2 + g.numel() * 4, f"[1024, 80, 256] features + [1024, 80, {n_fft // 2 + 1}, {frames}] magnitude gradients -> feature gradients "
             f"(2 reads + 1 write); torch autograd forward + backward through torch.stft(...).abs() on the same device: {ref_ms:.4f} ms (eager, includes its forward)")


if __name__ == "__main__":
    main()
