"""Throughput of the dataset driver mirror (preprocess/process_dataset.py) end to end: files on disk -> decode threads -> ragged
GPU batches -> {"mel"} payloads on disk.  One JSON line.

    python tools/bench_dataset_driver.py [--files 2000] [--seconds 10] [--decode-threads 8]

Synthetic 16-bit PCM .wav files (what LibriSpeech-style corpora hold once decoded) are written to a scratch directory first; the
timed region is ShardRunner.run over them, i.e. what `python preprocess/process_dataset.py --mel_only` does per GPU process."""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def _proc(rank, n_procs, files, a, decode_threads, start, done):
    """One of several driver processes on the same GPU: contiguous shard (process_dataset.py:256-259), warm-up, common start."""
    sys.path.insert(0, ROOT)
    import torch as _torch
    from audio_calm_b200.preprocess import process_dataset as pd
    from audio_calm_b200.sharding import contiguous_shard
    _torch.set_num_threads(1)
    mine = [files[i] for i in contiguous_shard(len(files), rank, n_procs)]
    warm = types.SimpleNamespace(**{**vars(a), "out_dir": a.out_dir + f"_warm{rank}"})
    pd.ShardRunner(warm, 0, decode_threads=decode_threads).run(mine[:16])
    shutil.rmtree(warm.out_dir, ignore_errors=True)
    runner = pd.ShardRunner(a, 0, decode_threads=decode_threads)
    start.wait()
    runner.run(mine)
    done.put((rank, len(mine), len(runner.errors)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--files", type=int, default=2000)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--decode-threads", type=int, default=8)
    ap.add_argument("--procs", type=int, default=1, help="driver processes sharing the GPU (--procs_per_gpu of the driver)")
    args = ap.parse_args()
    from scipy.io import wavfile
    from audio_calm_b200.preprocess import process_dataset as pd
    from oracle.logmel_oracle import hash_noise          # deterministic input only (a tool, not the product path)
    root = tempfile.mkdtemp(prefix="acb_driver_")
    try:
        in_dir, out_dir = os.path.join(root, "in"), os.path.join(root, "out")
        n = int(args.seconds * 16000)
        rng = np.random.default_rng(0)
        base = (hash_noise(n + 4096, 5) * 0.8 * 32767).astype(np.int16)
        for i in range(args.files):
            d = os.path.join(in_dir, f"spk{i % 40:03d}", f"chap{i % 7}")
            os.makedirs(d, exist_ok=True)
            off = int(rng.integers(0, 4096))
            wavfile.write(os.path.join(d, f"utt{i:06d}.wav"), 16000, base[off:off + n - int(rng.integers(0, n // 2))])
        files = pd.scan_files(in_dir)
        audio_s = sum(os.path.getsize(f) - 44 for f in files) / 2 / 16000
        a = types.SimpleNamespace(dataset_name="librispeech", in_dir=in_dir, out_dir=out_dir, vae_ckpt=None, mel_only=True, cv_tsv=None,
                                  num_gpus=1, workers_per_gpu=args.decode_threads, force=False)
        torch.set_num_threads(1)
        if args.procs > 1:
            import multiprocessing as mp
            ctx = mp.get_context("spawn")
            start, done = ctx.Barrier(args.procs + 1), ctx.Queue()
            ps = [ctx.Process(target=_proc, args=(r, args.procs, files, a, args.decode_threads, start, done)) for r in range(args.procs)]
            for p in ps:
                p.start()
            start.wait()
            t0 = time.perf_counter()
            res = [done.get() for _ in ps]
            dt = time.perf_counter() - t0
            for p in ps:
                p.join()
            written = sum(len(fs) for _, _, fs in os.walk(out_dir))
            print(json.dumps({"tool": "bench_dataset_driver", "files": len(files), "written": written, "errors": sum(r[2] for r in res),
                              "audio_hours": audio_s / 3600, "seconds": dt, "files_per_s": len(files) / dt,
                              "audio_hours_per_s": audio_s / 3600 / dt, "procs": args.procs, "decode_threads": args.decode_threads,
                              "host_cpus": os.cpu_count(), "payload": '{"mel": FloatTensor[80, T4]} per file, torch.save',
                              "note": "several driver processes on ONE GPU (--procs_per_gpu), contiguous shards, timed from a common start to the last one done"}))
            return
        runner = pd.ShardRunner(a, 0, decode_threads=args.decode_threads)
        runner.run(files[:32])                              # warm-up: kernels, allocator, thread pools
        shutil.rmtree(out_dir, ignore_errors=True)
        runner = pd.ShardRunner(a, 0, decode_threads=args.decode_threads)
        t0 = time.perf_counter()
        runner.run(files)
        dt = time.perf_counter() - t0
        written = sum(len(fs) for _, _, fs in os.walk(out_dir))
        print(json.dumps({"tool": "bench_dataset_driver", "files": len(files), "written": written, "errors": len(runner.errors),
                          "audio_hours": audio_s / 3600, "seconds": dt, "files_per_s": len(files) / dt,
                          "audio_hours_per_s": audio_s / 3600 / dt, "decode_threads": args.decode_threads,
                          "host_cpus": os.cpu_count(), "payload": '{"mel": FloatTensor[80, T4]} per file, torch.save',
                          "note": "end to end per GPU process: scipy/torchaudio decode -> int16 H2D -> peak + fused log-mel (+ pad-to-4) -> D2H -> torch.save"}))
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
