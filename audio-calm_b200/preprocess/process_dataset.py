"""Drop-in for the reference's ``preprocess/process_dataset.py`` (same CLI, same output tree, same payloads).

Reference flow (``process_dataset.py:75-226``): ``num_gpus * workers_per_gpu`` spawned processes, each looping over its chunk of
files ONE CLIP AT A TIME: ``torchaudio.load`` -> resample -> ``process_audio_chunk`` (CPU) -> H2D -> ``MelExtractor`` (about ten
library launches) -> reflect-pad T to a multiple of 4 -> ``torch.save({"mel": [80, T4]})``.

Here: ONE process per GPU.  ``workers_per_gpu`` decode threads feed clips to the process; clips are packed into ragged batches
and every batch is three launches (``acb_peak_abs`` + the fused log-mel kernel with peak normalisation and pad-to-4 fused in)
followed by one D2H copy; writer threads ``torch.save`` the per-clip payloads.  Kept from the reference: the flags
(``:230-238``), the mirrored output tree and ``<file_id>.pt`` names (``:113-122``), skip-if-exists unless ``--force``
(``:125-130``), the ``{"mel": FloatTensor[80, T4]}`` / ``{"latent", "vae_path"}`` payloads (``:152-168``), transcript files
(``:170-189, 209-214``), contiguous ``ceil(N / procs)`` sharding (``:256-259``) and the ``None`` end-of-worker message
(``:98-100, 217``).  Unlike the reference (``:197-202``), per-file errors -- decode, kernel AND save failures -- are counted and reported,
not swallowed; a clip's transcript line is written only after its payload has been saved, as in the reference (``:152-189``).

    python preprocess/process_dataset.py --dataset_name librispeech --in_dir IN --out_dir OUT --mel_only [--num_gpus N] [--force]
"""
from __future__ import annotations

import argparse
import csv
import os
import queue as queue_mod
import sys
import time
from concurrent.futures import Future, ThreadPoolExecutor
from typing import Callable, Dict, List, NamedTuple, Optional, Sequence, Tuple

import torch

try:
    from ..frontend import LogMelFrontend, pack_clips
    from .._lib import check as _lib_check
    from ..sharding import contiguous_shard
    from .core import load_vae, process_audio_chunk  # noqa: F401
except ImportError:  # run as a script / as top-level `preprocess.process_dataset`
    _root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if _root not in sys.path:
        sys.path.insert(0, _root)
    from audio_calm_b200.frontend import LogMelFrontend, pack_clips
    from audio_calm_b200._lib import check as _lib_check
    from audio_calm_b200.sharding import contiguous_shard
    from audio_calm_b200.preprocess.core import load_vae, process_audio_chunk  # noqa: F401

AUDIO_EXTENSIONS = {".wav", ".flac", ".mp3"}   # process_dataset.py:66
TARGET_SR = 16000
PAD_TO = 4                                      # process_dataset.py:147
REPORT_BATCH = 100                              # process_dataset.py:104


# ------------------------------------------------------------------------------------------- host-side bookkeeping
def scan_files(root_dir: str) -> List[str]:
    """Every audio file under ``root_dir`` in ``os.walk`` order (process_dataset.py:60-73)."""
    files = []
    for root, _, filenames in os.walk(root_dir):
        for f in filenames:
            if os.path.splitext(f)[1].lower() in AUDIO_EXTENSIONS:
                files.append(os.path.join(root, f))
    return files


def get_common_voice_map(tsv_path: Optional[str]) -> Dict[str, str]:
    """CommonVoice ``path -> sentence`` (process_dataset.py:31-41)."""
    mapping: Dict[str, str] = {}
    if not tsv_path or not os.path.exists(tsv_path):
        return mapping
    with open(tsv_path, "r", encoding="utf-8") as f:
        for row in csv.DictReader(f, delimiter="\t"):
            mapping[row["path"]] = row["sentence"]
    return mapping


def output_path(wav_path: str, args) -> Tuple[str, str, str]:
    """(save_dir, file_id, save_path): flat for commonvoice, mirrored tree otherwise (process_dataset.py:113-122)."""
    file_id = os.path.splitext(os.path.basename(wav_path))[0]
    if args.dataset_name == "commonvoice":
        save_dir = args.out_dir
    else:
        save_dir = os.path.join(args.out_dir, os.path.relpath(os.path.dirname(wav_path), args.in_dir))
    return save_dir, file_id, os.path.join(save_dir, f"{file_id}.pt")


def transcript_for(wav_path: str, args, cv_mapping: Dict[str, str]) -> Optional[str]:
    """Transcript lookup per dataset flavour (process_dataset.py:43-58, 170-178)."""
    if args.dataset_name == "libritts":
        txt = wav_path.replace(".wav", ".normalized.txt")
        if os.path.exists(txt):
            with open(txt, "r", encoding="utf-8") as f:
                return f.read().strip()
    elif args.dataset_name == "librispeech":
        folder = os.path.dirname(wav_path)
        file_id = os.path.splitext(os.path.basename(wav_path))[0]
        try:
            trans = next((f for f in os.listdir(folder) if f.endswith(".trans.txt")), None)
            if trans:
                with open(os.path.join(folder, trans), "r", encoding="utf-8") as f:
                    for line in f:
                        if line.startswith(file_id):
                            return line.strip().split(" ", 1)[1]
        except (OSError, IndexError):
            return None
    elif args.dataset_name == "commonvoice":
        return cv_mapping.get(os.path.basename(wav_path))
    return None


def load_audio(path: str) -> torch.Tensor:
    """Decode (+ resample to 16 kHz) on the host: ``[C, L]`` (process_dataset.py:135-137).  I/O, not on the hot path.

    16-bit files that already are 16 kHz come back as **int16**: the driver ships them as PCM (half the PCIe bytes) and widens them
    on the device to ``x / 32768`` -- exactly what ``torchaudio.load`` (normalize=True) would have produced.  Everything else is float32."""
    try:
        import torchaudio
        wav, sr = torchaudio.load(path, normalize=False)
    except Exception:  # noqa: BLE001 - torchaudio without a decoding backend: plain PCM .wav through scipy
        from scipy.io import wavfile
        import numpy as np
        sr, data = wavfile.read(path)
        wav = torch.from_numpy(np.ascontiguousarray(np.atleast_2d(data.T if data.ndim == 2 else data)))
    if wav.dtype == torch.int16 and sr == TARGET_SR:
        return wav.contiguous()
    if wav.dtype == torch.int16:
        wav = wav.to(torch.float32) / 32768.0
    elif wav.dtype == torch.int32:
        wav = wav.to(torch.float32) / 2147483648.0
    elif wav.dtype == torch.uint8:
        wav = (wav.to(torch.float32) - 128.0) / 128.0
    else:
        wav = wav.to(torch.float32)
    if sr != TARGET_SR:
        import torchaudio
        wav = torchaudio.transforms.Resample(sr, TARGET_SR)(wav)
    return wav


# ------------------------------------------------------------------------------------------- the per-GPU engine
class Clip(NamedTuple):
    wav_path: str
    save_dir: str
    file_id: str
    save_path: str
    text: Optional[str]
    wav: torch.Tensor


class ShardRunner:
    """Processes one shard of files on one GPU in ragged batches."""

    def __init__(self, args, gpu_id: int, cv_mapping: Optional[Dict[str, str]] = None, load_fn: Callable = load_audio,
                 batch_samples: int = 64 * 30 * TARGET_SR, decode_threads: int = 4, report: Optional[Callable[[int], None]] = None):
        self.args, self.cv_mapping, self.load_fn = args, cv_mapping or {}, load_fn
        self.device = torch.device("cuda", gpu_id)
        self.batch_samples, self.decode_threads = int(batch_samples), max(1, int(decode_threads))
        self.report = report or (lambda n: None)
        self.fe = LogMelFrontend(self.device)
        self.vae = None
        if not args.mel_only:
            self.vae = load_vae(args.vae_ckpt, self.device)  # needs the reference's models/ on sys.path (out of scope here)
        self.errors: List[Tuple[str, str]] = []
        self.trans_buffer: Dict[str, List[str]] = {}
        self.done = 0

    # -- one ragged batch: peak -> fused log-mel (+ peak norm, + pad-to-4) -> D2H -> save
    def _flush(self, items: List[Clip], writer: ThreadPoolExecutor) -> None:
        if not items:
            return
        lib = self.fe._lib
        with torch.inference_mode(), torch.cuda.device(self.device):
            # Peak normalisation is fused into the log-mel kernel (per-clip gain from acb_peak_abs).  Multi-channel clips are
            # mixed down first (acb_mixdown_peak: channel mean only) and then take the same fused route.
            clips = []
            for it in items:
                w = it.wav.to(self.device, non_blocking=True)
                if w.dtype == torch.int16:
                    w = self.fe.pcm16_to_float(w)          # 16-bit PCM transport, widened on the device
                if w.shape[0] == 1:
                    clips.append(w[0])
                else:
                    w = w.contiguous()
                    mono = torch.empty(w.shape[1], dtype=torch.float32, device=self.device)
                    _lib_check(lib.acb_mixdown_peak(w.data_ptr(), int(w.shape[0]), int(w.shape[1]), mono.data_ptr(), None,
                                                    torch.cuda.current_stream(self.device).cuda_stream), "acb_mixdown_peak")
                    clips.append(mono)
            batch = pack_clips(clips, self.device)
            peak = self.fe.peak_abs_ragged(batch)
            feats, frames = self.fe.forward_ragged(batch, pad_multiple=PAD_TO, peak=peak)      # [B, 80, Tmax], frames[B] = T4
            frames_h = frames.cpu().tolist()
            if self.vae is None:
                host = feats.cpu()
                for i, it in enumerate(items):
                    mel = host[i, :, :frames_h[i]].clone()                                     # logical [80, T4] float32
                    self._inflight.append((writer.submit(self._save, it.save_dir, it.save_path, {"mel": mel}), it))
            else:
                for i, it in enumerate(items):
                    mu, _ = self.vae.encode(feats[i:i + 1, :, :frames_h[i]])                   # process_dataset.py:159-163
                    payload = {"latent": mu.squeeze(0).cpu(), "vae_path": self.args.vae_ckpt}
                    self._inflight.append((writer.submit(self._save, it.save_dir, it.save_path, payload), it))

    @staticmethod
    def _save(save_dir: str, save_path: str, payload: dict) -> None:
        os.makedirs(save_dir, exist_ok=True)
        torch.save(payload, save_path)

    def _collect(self, wait: bool) -> None:
        """Harvest finished saves: a failed save (disk full, permissions ...) becomes a reported error; a clip's transcript line
        is buffered only once its payload is on disk (process_dataset.py:152-189 writes the entry after torch.save)."""
        keep: List[Tuple[Future, Clip]] = []
        for fut, it in self._inflight:
            if not wait and not fut.done():
                keep.append((fut, it))
                continue
            try:
                fut.result()
            except Exception as e:  # noqa: BLE001
                self.errors.append((it.wav_path, f"save failed: {type(e).__name__}: {e}"))
            else:
                if it.text:
                    fname = ("commonvoice.trans.txt" if self.args.dataset_name == "commonvoice"
                             else f"{os.path.basename(it.save_dir)}.trans.txt")
                    self.trans_buffer.setdefault(os.path.join(it.save_dir, fname), []).append(f"{it.file_id} {it.text}")
            self._tick()
        self._inflight = keep

    def _tick(self, n: int = 1) -> None:
        self.done += n
        if self.done >= REPORT_BATCH:
            self.report(self.done)
            self.done = 0

    def run(self, file_list: Sequence[str]) -> None:
        args = self.args
        todo = []
        for wav_path in file_list:
            save_dir, file_id, save_path = output_path(wav_path, args)
            if os.path.exists(save_path) and not args.force:                                   # resume (process_dataset.py:125-130)
                self._tick()
                continue
            todo.append((wav_path, save_dir, file_id, save_path))

        def decode(job):
            wav_path = job[0]
            try:
                return job, self.load_fn(wav_path), None
            except Exception as e:  # noqa: BLE001
                return job, None, f"{type(e).__name__}: {e}"

        pending: List[Clip] = []
        pending_samples = 0
        self._inflight: List[Tuple[Future, Clip]] = []
        with ThreadPoolExecutor(self.decode_threads) as decoders, ThreadPoolExecutor(2) as writer:
            for job, wav, err in decoders.map(decode, todo):
                wav_path, save_dir, file_id, save_path = job
                if err is None and (wav.dim() != 2 or wav.shape[-1] <= self.fe.n_fft // 2):
                    err = f"clip of {tuple(wav.shape)} samples is too short for reflect padding"
                if err is not None:
                    self.errors.append((wav_path, err))
                    self._tick()
                    continue
                if pending and pending_samples + wav.shape[-1] > self.batch_samples:
                    self._flush_safe(pending, writer)
                    pending, pending_samples = [], 0
                    self._collect(wait=False)
                pending.append(Clip(wav_path, save_dir, file_id, save_path, transcript_for(wav_path, args, self.cv_mapping), wav))
                pending_samples += int(wav.shape[-1])
            self._flush_safe(pending, writer)
            self._collect(wait=True)
        if self.done:
            self.report(self.done)
            self.done = 0
        for path, lines in self.trans_buffer.items():                                          # process_dataset.py:209-214
            os.makedirs(os.path.dirname(path), exist_ok=True)
            with open(path, "a", encoding="utf-8") as f:
                f.writelines(line + "\n" for line in lines)

    def _flush_safe(self, items: List[Clip], writer) -> None:
        n_before = len(self._inflight)
        try:
            self._flush(items, writer)
        except Exception as e:  # noqa: BLE001 - report, keep the shard going
            submitted = {id(it) for _, it in self._inflight[n_before:]}
            for it in items:
                if id(it) not in submitted:                                                    # clips whose save was never queued
                    self.errors.append((it.wav_path, f"{type(e).__name__}: {e}"))
                    self._tick()


def worker_process(rank: int, gpu_id: int, file_list: Sequence[str], args, cv_mapping, queue) -> None:
    """One process per GPU; messages on ``queue``: ints = files finished, ("errors", [...]), then None (process_dataset.py:75-217)."""
    torch.set_num_threads(1)                                                                   # process_dataset.py:86
    try:
        runner = ShardRunner(args, gpu_id, cv_mapping, decode_threads=args.workers_per_gpu, report=queue.put)
    except Exception as e:  # noqa: BLE001 - init failure: same protocol as the reference (:98-100), plus the reason
        queue.put(("errors", [(f"<worker {rank} init>", f"{type(e).__name__}: {e}")]))
        queue.put(None)
        return
    runner.run(file_list)
    if runner.errors:
        queue.put(("errors", runner.errors))
    queue.put(None)


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser()
    p.add_argument("--dataset_name", type=str, required=True, help="libritts | librispeech | commonvoice")
    p.add_argument("--in_dir", type=str, required=True)
    p.add_argument("--out_dir", type=str, required=True)
    p.add_argument("--vae_ckpt", type=str, default=None)
    p.add_argument("--mel_only", action="store_true")
    p.add_argument("--cv_tsv", type=str, default=None)
    p.add_argument("--num_gpus", type=int, default=torch.cuda.device_count())
    p.add_argument("--workers_per_gpu", type=int, default=4, help="decode threads per GPU process (the reference spawns this many processes)")
    p.add_argument("--procs_per_gpu", type=int, default=1,
                   help="processes per GPU, each with --workers_per_gpu decode threads (decoding and torch.save hold the interpreter lock: "
                        "more processes scale the host side like the reference's num_gpus * workers_per_gpu processes do, rank %% num_gpus "
                        "picks the device, process_dataset.py:256-275)")
    p.add_argument("--force", action="store_true")
    return p


def main(argv: Optional[Sequence[str]] = None) -> int:
    import multiprocessing as mp
    args = build_parser().parse_args(argv)
    if not args.mel_only and args.vae_ckpt is None:
        print("Error: extracting latents (without --mel_only) needs --vae_ckpt")
        return 2
    if args.num_gpus < 1:
        raise RuntimeError("process_dataset (B200 build) needs at least one CUDA device; there is no CPU fallback")
    files = scan_files(args.in_dir)
    print(f"Found {len(files)} files in {args.in_dir}", flush=True)
    if not files:
        return 0
    cv_mapping = get_common_voice_map(args.cv_tsv) if args.dataset_name == "commonvoice" else {}
    ctx = mp.get_context("spawn")                                                             # process_dataset.py:267
    queue = ctx.Manager().Queue()
    procs = []
    n_procs = args.num_gpus * max(1, int(getattr(args, "procs_per_gpu", 1)))
    for rank in range(n_procs):                                                                # contiguous ceil(N / procs) chunks (:256-259)
        shard = contiguous_shard(len(files), rank, n_procs)
        if len(shard) == 0:
            continue
        p = ctx.Process(target=worker_process, args=(rank, rank % args.num_gpus, [files[i] for i in shard], args, cv_mapping, queue))
        p.start()
        procs.append(p)
    done, finished, errors, t0 = 0, 0, [], time.time()
    while finished < len(procs):
        try:
            msg = queue.get(timeout=0.5)
        except queue_mod.Empty:
            if not any(p.is_alive() for p in procs) and queue.empty():                         # process_dataset.py:302-305
                break
            continue
        if msg is None:
            finished += 1
        elif isinstance(msg, int):
            done += msg
            sys.stdout.write(f"\rProcessing {os.path.basename(args.in_dir.rstrip('/'))}: {done}/{len(files)} "
                             f"[{done / (time.time() - t0 + 1e-5):.1f} file/s]")
            sys.stdout.flush()
        elif isinstance(msg, tuple) and msg[0] == "errors":
            errors.extend(msg[1])
    print(f"\nDone: {done}/{len(files)} files in {time.time() - t0:.1f} s, {len(errors)} error(s)")
    for path, err in errors[:20]:
        print(f"  {path}: {err}")
    for p in procs:
        p.join()
    return 0 if not errors else 1


if __name__ == "__main__":
    sys.exit(main())
