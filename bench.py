"""Benchmark of the log-mel front-end hot path (BASELINE.json metric: log-mel audio-hours/sec; % of HBM roofline).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA path through the C ABI)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (torch.stft) on the host cores

Workload at every N: BASELINE.json configs[1] -- a batch of 256 x 30 s synthetic 16 kHz clips per GPU ->
scalar-normalised log-mel fp32 [256, 80, 1876].  A "step" is one pass of the fused kernel over one batch; at N > 1 every
rank processes its own batch (utterances shard with no data-path collective => weak scaling).  Prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

SAMPLE_RATE = 16000
METRIC = "log-mel audio-hours/sec"
UNIT = "audio-hours/s"


def synth_batch(n_clips: int, length: int, device, seed: int = 1234) -> torch.Tensor:
    """N(0, 0.1^2) x slow envelope 0.25 + 0.75 sin^2(2 pi 0.7 t), clipped to +-1, last 5 % exact zeros."""
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(n_clips, length, device=device, generator=g) * 0.1
    t = torch.arange(length, device=device, dtype=torch.float32) / SAMPLE_RATE
    x = (x * (0.25 + 0.75 * torch.sin(2 * np.pi * 0.7 * t) ** 2)).clamp_(-1.0, 1.0)
    x[:, length - length // 20:] = 0.0
    return x


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                try:
                    mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    mask = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.0005)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": int(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:  # noqa: BLE001
        return None


def cpu_reference_rate(n_clips: int, length: int, min_seconds: float, threads: int):
    """Time the reference's CPU path (oracle/ref_torch_port.py: torchaudio MelSpectrogram / torch.stft + log(clamp) +
    (x - mean)/std) on a bounded sample with `threads` host threads.  Returns (audio-hours/s, iterations, seconds)."""
    from oracle.ref_torch_port import RefMelExtractor, normalise
    torch.set_num_threads(threads)
    ext = RefMelExtractor().eval()
    x = synth_batch(n_clips, length, "cpu")
    with torch.inference_mode():
        normalise(ext(x[:2]))
        it, t0 = 0, time.perf_counter()
        while True:
            normalise(ext(x))
            it += 1
            dt = time.perf_counter() - t0
            if dt >= min_seconds and it >= 2:
                break
    return n_clips * length / SAMPLE_RATE * it / dt / 3600.0, it, dt


def run_reference(args) -> None:
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.ref_torch_port import RefMelExtractor, normalise
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    n_clips, length = args.ref_clips, args.seconds * SAMPLE_RATE
    ext = RefMelExtractor().eval()
    x = synth_batch(n_clips, length, "cpu")
    with torch.inference_mode():
        for _ in range(args.warmup):
            normalise(ext(x))
        t0 = time.perf_counter()
        for _ in range(args.steps):
            normalise(ext(x))
        dt = time.perf_counter() - t0
    value = n_clips * args.seconds * args.steps / dt / 3600.0
    sample = f"{n_clips} x {args.seconds} s clips per step (of the 256-clip batch), batched call, torch.set_num_threads({threads})"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config2: batch 256 x 30 s 16 kHz clips -> normalised log-mel fp32 [80 x 1876] (bounded CPU sample)",
                   "clips_per_step": n_clips, "clip_seconds": args.seconds},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="clips per GPU per step")
    ap.add_argument("--seconds", type=int, default=30, help="clip length")
    ap.add_argument("--ref-clips", type=int, default=16, help="clips per step of the CPU reference arm")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-whisper", action="store_true", help="skip the extra line of the tensor-core route (Whisper-style preset)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    if args.impl == "reference":
        run_reference(args)
        return

    import audio_calm_b200 as acb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":     # keeps "NCCL version ..." off stdout: rank 0 prints ONE JSON line
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)

    warm = max(args.warmup, 3)                       # timing rule: at least 3 warm-up steps
    B, L = args.batch, args.seconds * SAMPLE_RATE
    fe = acb.LogMelFrontend(device)
    T = fe.frames_for_length(L)
    affine = (acb.MEL_MEAN_DEFAULT, acb.MEL_STD_DEFAULT)
    x = synth_batch(B, L, device, seed=1234 + rank)
    out = torch.empty((B, fe.n_mels, T), dtype=torch.float32, device=device)
    stream = torch.cuda.current_stream(device)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ------------------------------------------------------------------ device-resident timing
    for _ in range(warm):
        fe.forward(x, affine=affine, out=out)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = fe.launches
    t_start.record(stream)
    for i in range(args.steps):
        ev[i][0].record(stream)
        fe.forward(x, affine=affine, out=out)
        ev[i][1].record(stream)
    t_end.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = fe.launches - launches0
    total_ms = t_start.elapsed_time(t_end)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    if dist is not None:
        tt = torch.tensor([total_ms], dtype=torch.float64, device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    audio_s_per_step = B * args.seconds * world
    value = audio_s_per_step * args.steps / (total_ms * 1e-3) / 3600.0

    # roofline: algorithmic bytes = every sample read once + every output written once
    alg_bytes = B * (4 * L + 4 * fe.n_mels * T)
    peak, peak_src = measured_peak()
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic("config2"), "kernel": "logmel_fused_kernel", "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "frames_per_s": B * T / (kernel_ms * 1e-3)}

    # ------------------------------------------------------------------ end to end through the public host-buffer API
    e2e = None
    if not args.no_e2e:
        x_host = torch.empty((B, L), dtype=torch.float32, pin_memory=True)
        x_host.copy_(x)
        out_host = torch.empty((B, fe.n_mels, T), dtype=torch.float32, pin_memory=True)
        staging = (torch.empty_like(x), out)
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            fe.forward_host(x_host, out_host, affine=affine, n_chunks=32, staging=staging)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fe.forward_host(x_host, out_host, affine=affine, n_chunks=32, staging=staging)   # synchronises: result is on the host
        torch.cuda.synchronize(device)
        dt = time.perf_counter() - t0
        if dist is not None:
            tt = torch.tensor([dt], dtype=torch.float64, device=device)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e2e = {"value": audio_s_per_step * e2e_steps / dt / 3600.0, "unit": UNIT, "h2d_bytes_per_step": int(x_host.numel() * 4),
               "d2h_bytes_per_step": int(out_host.numel() * 4), "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
               "api": "LogMelFrontend.forward_host -> acb_logmel_forward_host (pinned host in/out, 32 chunks, 3 streams)"}
        # extra, not the headline: the same step fed with 16-bit PCM (what audio files hold), widened on the device -- half the H2D bytes
        pcm_host = torch.empty((B, L), dtype=torch.int16, pin_memory=True)
        pcm_host.copy_((x * 32767.0).to(torch.int16))
        staging = staging + (torch.empty((B, L), dtype=torch.int16, device=device),)
        for _ in range(2):
            fe.forward_host(pcm_host, out_host, affine=affine, n_chunks=32, staging=staging)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fe.forward_host(pcm_host, out_host, affine=affine, n_chunks=32, staging=staging)
        torch.cuda.synchronize(device)
        dt16 = time.perf_counter() - t0
        if dist is not None:
            tt = torch.tensor([dt16], dtype=torch.float64, device=device)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt16 = float(tt.item())
        e2e["pcm16_input"] = {"value": audio_s_per_step * e2e_steps / dt16 / 3600.0, "unit": UNIT, "h2d_bytes_per_step": int(pcm_host.numel() * 2),
                              "d2h_bytes_per_step": int(out_host.numel() * 4), "ms_per_step": dt16 / e2e_steps * 1e3,
                              "note": "extension: int16 PCM host input (bit-identical features); the fp32 figure above is the comparable one"}
        del x_host, out_host, staging, pcm_host

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, it, dt = cpu_reference_rate(args.ref_clips, L, args.cpu_seconds, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{args.ref_clips} x {args.seconds} s clips x {it} iterations ({dt:.1f} s), oracle/ref_torch_port.py "
                         f"(torchaudio MelSpectrogram + log(clamp) + normalise), torch.set_num_threads({threads})"}

    # ------------------------------------------------------------------ extra: the tensor-core route (Whisper-style preset)
    whisper = None
    if rank == 0 and not args.no_whisper:
        try:
            wfe = acb.WhisperLogMel(device)
            wout = torch.empty((B, 80, wfe.frames_for_length(L)), dtype=torch.float32, device=device)
            for _ in range(3):
                wfe.forward(x, out=wout)
            torch.cuda.synchronize(device)
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record(stream)
            for _ in range(20):
                wfe.forward(x, out=wout)
            w1.record(stream)
            torch.cuda.synchronize(device)
            wms = w0.elapsed_time(w1) / 20
            walg = B * (4 * L + 4 * 80 * wout.shape[2])
            whisper = {"value": B * args.seconds / (wms * 1e-3) / 3600.0, "unit": UNIT, "ms_per_step": wms,
                       "frames_per_s": B * wout.shape[2] / (wms * 1e-3), "roofline_frac": walg / (wms * 1e-3) / 1e9 / peak,
                       "kernel": "dftgemm_logmel_kernel (tcgen05 DFT-GEMM, n_fft 400 / hop 160) + dftgemm_floor_kernel (tiles below the floor only)",
                       "note": "extension preset (north-star wording); tools/bench_whisper.py prints its full line"}
            del wfe, wout
        except Exception as e:  # noqa: BLE001
            whisper = {"unavailable": str(e)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config2: batch {B} x {args.seconds} s 16 kHz clips per GPU -> normalised log-mel fp32 [{B}, 80, {T}]",
                       "clips_per_gpu": B, "clip_seconds": args.seconds, "n_fft": 1024, "hop": 256, "n_mels": 80,
                       "sharding": "by utterance, no data-path collective",
                       "l2": f"inputs {B * L * 4 / 1e6:.1f} MB + outputs {B * 80 * T * 4 / 1e6:.1f} MB per step exceed the 126 MB L2; no flush"},
            "audio_seconds_per_s": value * 3600.0,
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "whisper_preset": whisper,
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
