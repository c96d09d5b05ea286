"""Benchmark of the tensor-core route (Whisper-style preset: n_fft 400, hop 160, DFT-as-GEMM on tcgen05, csrc/acb_dftgemm.cu).

    python tools/bench_whisper.py [--batch 256] [--steps 50] [--warmup 3] [--no-cpu-baseline]

Workload: 256 x 30 s synthetic 16 kHz clips per GPU -> WhisperFeatureExtractor-compatible features fp32 [256, 80, 3000]
(both launches: the tcgen05 kernel and the sparse dynamic-range floor pass).  Prints ONE JSON line shaped like bench.py's.
Algorithmic bytes per frame: 160 samples x 4 B + 80 x 4 B = 960 B; GEMM flops per frame: 4 GEMMs x 3 split terms x 2 x 112 x 112.
The CPU baseline is the unmodified transformers.WhisperFeatureExtractor (numpy path) on the host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

from bench import ClockSampler, measured_peak, ncu_record as ncu_traffic, synth_batch  # noqa: E402

SAMPLE_RATE = 16000


def tensor_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:  # noqa: BLE001
        return 1416.5, "fallback (1416.5 TFLOP/s sustained bf16)"


def measure(device, batch: int, seconds: int, steps: int, warmup: int):
    import audio_calm_b200 as acb
    fe = acb.WhisperLogMel(device)
    L = seconds * SAMPLE_RATE
    x = synth_batch(batch, L, device)
    T = fe.frames_for_length(L)
    out = torch.empty((batch, fe.n_mels, T), dtype=torch.float32, device=device)
    for _ in range(max(warmup, 3)):
        fe.forward(x, out=out)
    fe.forward(x, out=out, check=True)
    torch.cuda.synchronize(device)
    stream = torch.cuda.current_stream(device)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    n0 = fe.launches
    for a, b in ev:
        a.record(stream)
        fe.forward(x, out=out)
        b.record(stream)
    torch.cuda.synchronize(device)
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    return fe, x, out, T, ms, fe.launches - n0


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--seconds", type=int, default=30)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise RuntimeError("needs a CUDA device; there is no CPU fallback")
    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)
    sampler = ClockSampler(0)
    sampler.start()
    fe, x, out, T, ms, launches = measure(device, args.batch, args.seconds, args.steps, args.warmup)
    clocks = sampler.stop()
    B, L = args.batch, args.seconds * SAMPLE_RATE
    value = B * args.seconds / (ms * 1e-3) / 3600.0
    alg = B * (4 * L + 4 * fe.n_mels * T)
    peak, src = measured_peak()
    ach = alg / (ms * 1e-3) / 1e9
    flops = B * (-(-T // 128) * 128) * 4 * 3 * 2 * 112 * 112
    tpeak, tsrc = tensor_peak()
    # end to end through the public host-buffer API (pinned host in/out, chunked H2D / kernels / D2H on three streams)
    x_host = torch.empty((B, L), dtype=torch.float32, pin_memory=True)
    x_host.copy_(x)
    out_host = torch.empty((B, fe.n_mels, T), dtype=torch.float32, pin_memory=True)
    for _ in range(2):
        fe.forward_host(x_host, out_host, n_chunks=16)
    t0 = time.perf_counter()
    e2e_steps = 10
    for _ in range(e2e_steps):
        fe.forward_host(x_host, out_host, n_chunks=16)
    dt = (time.perf_counter() - t0) / e2e_steps
    e2e = {"value": B * args.seconds / dt / 3600.0, "unit": "audio-hours/s", "h2d_bytes_per_step": int(x_host.numel() * 4),
           "d2h_bytes_per_step": int(out_host.numel() * 4), "ms_per_step": dt * 1e3, "steps": e2e_steps,
           "api": "WhisperLogMel.forward_host (pinned host in/out, 16 chunks, 3 streams)",
           "max_abs_diff_vs_device_path": float((out_host - out.cpu()).abs().max())}
    cpu = None
    if not args.no_cpu_baseline:
        try:
            from transformers import WhisperFeatureExtractor
            hf = WhisperFeatureExtractor()
            clips = [c for c in x[:4].cpu().numpy()]
            hf(clips[:1], sampling_rate=SAMPLE_RATE, return_tensors="np")
            it, t0 = 0, time.perf_counter()
            while True:
                ref = hf(clips, sampling_rate=SAMPLE_RATE, return_tensors="np")["input_features"]
                it += 1
                dt = time.perf_counter() - t0
                if dt >= args.cpu_seconds and it >= 2:
                    break
            err = float(np.abs(ref - out[:4].cpu().numpy()).max())
            cpu = {"value": len(clips) * args.seconds * it / dt / 3600.0, "unit": "audio-hours/s", "cores": 1, "kind": "reference",
                   "sample": f"{len(clips)} x {args.seconds} s clips x {it} iterations ({dt:.1f} s), transformers.WhisperFeatureExtractor (numpy path, single thread)",
                   "max_abs_diff_vs_gpu": err}
        except Exception as e:  # noqa: BLE001
            cpu = {"unavailable": str(e)[:200]}
    print(json.dumps({
        "metric": "log-mel audio-hours/sec", "value": value, "unit": "audio-hours/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (fp16 x 3 split operands, fp32 accumulate)",
        "data": "synthetic",
        "config": {"workload": f"whisper preset: batch {B} x {args.seconds} s 16 kHz clips -> log10-mel fp32 [{B}, 80, {T}] (max-8 floor, (x+4)/4)",
                   "n_fft": 400, "hop": 160, "n_mels": 80, "route": "DFT-as-GEMM on tcgen05 (4 real GEMMs 128x112x112 per 128 frames, 3 split terms)",
                   "l2": f"inputs {B * L * 4 / 1e6:.1f} MB + outputs {B * 80 * T * 4 / 1e6:.1f} MB per step exceed the 126 MB L2; no flush"},
        "frames_per_s": B * T / (ms * 1e-3), "clocks": clocks, "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": ncu_traffic("whisper"),
                     "kernel": "dftgemm_logmel_kernel + dftgemm_floor_kernel", "kernel_ms": ms, "algorithmic_bytes_per_launch": alg, "peak_source": src},
        "tensor": {"achieved": flops / (ms * 1e-3) / 1e12, "peak": tpeak, "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / tpeak,
                   "flops_per_step": flops, "peak_source": tsrc},
        "e2e": e2e, "cpu_baseline": cpu,
    }))


if __name__ == "__main__":
    main()
