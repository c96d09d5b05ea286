"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (``python oracle/gen_golden.py``): it imports
``/root/reference/preprocess/core.py`` (MelExtractor, process_audio_chunk) and restates, around those calls,
the few lines of ``preprocess/process_dataset.py:140-156`` (pad-to-4) and
``preprocess/compute_mel_stats.py:19-36`` (statistics loop) that cannot be imported as functions.
The reference tree does not exist on the GPU box, so the outputs are committed as fixtures.

Environment the vectors were produced with is recorded in ``tests/golden/MANIFEST.json``.
"""
from __future__ import annotations

import io
import json
import math
import os
import sys
import contextlib

sys.dont_write_bytecode = True
REF = os.environ.get("AUDIOCALM_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

import numpy as np
import torch
import torchaudio

from oracle.logmel_oracle import hash_noise, synth_clip  # deterministic inputs only

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main() -> None:
    torch.set_num_threads(1)
    from preprocess.core import MelExtractor, process_audio_chunk  # the reference itself

    os.makedirs(OUT, exist_ok=True)
    ext = MelExtractor().eval()
    window = ext.mel_transform.spectrogram.window.numpy().copy()
    fb = ext.mel_transform.mel_scale.fb.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "tables.npz"), window=window, fb=fb)

    def ref_pipeline(wav_cl: np.ndarray, peak_norm: bool = True, pad4: bool = True) -> np.ndarray:
        with torch.inference_mode():
            w = torch.from_numpy(np.ascontiguousarray(wav_cl))
            if peak_norm:
                w = process_audio_chunk(w)                              # preprocess/process_dataset.py:140
            mel = ext(w)                                                # :144
            if pad4 and mel.shape[-1] % 4 != 0:                         # :146-150
                pad_len = 4 - (mel.shape[-1] % 4)
                mel = torch.nn.functional.pad(mel, (0, pad_len), mode="reflect")
            return mel.squeeze(0).contiguous().numpy().copy()           # :155

    cases = {}
    # --- hash-noise clips through the full dataset pipeline (SURVEY.md §8c starter vectors) ---
    for n, seed in ((16000, 1), (40000, 2), (100001, 3)):
        cases[f"pipeline_noise_{n}_s{seed}"] = ref_pipeline(hash_noise(n, seed)[None])
    cases["raw_noise_16000_s1"] = ref_pipeline(hash_noise(16000, 1)[None], peak_norm=False, pad4=False)
    # --- speech-like synthetic clips (bench distribution), odd lengths, minimum lengths ---
    for n, seed in ((8000, 11), (24001, 12), (513, 13), (1024, 14), (777, 15), (1279, 16), (1280, 17)):
        cases[f"raw_synth_{n}_s{seed}"] = ref_pipeline(synth_clip(n, seed)[None], peak_norm=False, pad4=False)
    cases["pipeline_synth_24001_s12"] = ref_pipeline(synth_clip(24001, 12)[None])
    # --- tone: clamp behaviour (bins at the floor) ---
    t = np.arange(16000, dtype=np.float64) / 16000.0
    tone = (0.5 * np.sin(2 * np.pi * 440 * t) + 0.25 * np.sin(2 * np.pi * 3000 * t + 1.0)).astype(np.float32)
    cases["raw_tone_16000"] = ref_pipeline(tone[None], peak_norm=False, pad4=False)
    # --- silence: every value at the floor ---
    cases["raw_zeros_4000"] = ref_pipeline(np.zeros((1, 4000), np.float32), peak_norm=False, pad4=False)
    # --- batched input: [B, L] -> [B, 80, T] ---
    batch = np.stack([synth_clip(12000, 21), synth_clip(12000, 22), hash_noise(12000, 23)])
    with torch.inference_mode():
        cases["raw_batch3_12000"] = ext(torch.from_numpy(batch)).contiguous().numpy().copy()

    # --- process_audio_chunk on its own: stereo mix-down + peak normalisation; silent clip ---
    stereo = np.stack([hash_noise(5000, 31), synth_clip(5000, 32)])
    with torch.inference_mode():
        cases["chunk_stereo_5000"] = process_audio_chunk(torch.from_numpy(stereo)).numpy().copy()
        cases["chunk_mono_5000"] = process_audio_chunk(torch.from_numpy(synth_clip(5000, 33)[None])).numpy().copy()
        cases["chunk_zeros_100"] = process_audio_chunk(torch.zeros(1, 100)).numpy().copy()

    # --- normalisations ---
    mel1 = torch.from_numpy(cases["pipeline_noise_16000_s1"])[None]
    cases["norm_global_noise_16000_s1"] = ((mel1 - (-6.589515)) / 3.860679).squeeze(0).numpy().copy()   # modeling_vae.py:317-319
    mean = mel1.mean(dim=-1, keepdim=True)                                                          # eval_vae.py:80-82
    std = mel1.std(dim=-1, keepdim=True).clamp(min=1e-5)
    cases["norm_utt_noise_16000_s1"] = ((mel1 - mean) / std).squeeze(0).numpy().copy()

    np.savez_compressed(os.path.join(OUT, "cases.npz"), **cases)

    # --- statistics pass over the three saved "files" (compute_mel_stats.py:19-36) ---
    total_sum = 0.0
    total_sq_sum = 0.0
    total_count = 0
    files = [cases[f"pipeline_noise_{n}_s{s}"] for n, s in ((16000, 1), (40000, 2), (100001, 3))]
    for arr in files:
        mel = torch.from_numpy(arr).float()
        total_sum += mel.sum().item()
        total_sq_sum += (mel ** 2).sum().item()
        total_count += mel.numel()
    mean_s = total_sum / total_count
    var_s = max(total_sq_sum / total_count - mean_s * mean_s, 1e-8)
    std_s = math.sqrt(var_s)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        print(f"Global mel_mean: {mean_s:.6f}")
        print(f"Global mel_std:  {std_s:.6f}")
    cat = np.concatenate([f.astype(np.float64) for f in files], axis=1)
    stats = {
        "total_sum": total_sum, "total_sq_sum": total_sq_sum, "total_count": total_count,
        "mean": mean_s, "std": std_s, "printed": buf.getvalue().splitlines(),
        "per_bin_mean_0_40_79": [float(cat.mean(axis=1)[i]) for i in (0, 40, 79)],
        "per_bin_std_0_40_79": [float(cat.std(axis=1)[i]) for i in (0, 40, 79)],
    }

    # --- frame-count table (bit exact) and the error for too-short inputs ---
    frame_table = {}
    for L in (513, 514, 767, 768, 1023, 1024, 1025, 8000, 16000, 160000, 480000):
        with torch.inference_mode():
            frame_table[str(L)] = int(ext(torch.zeros(1, L)).shape[-1])
    short_error = {}
    for L in (1, 256, 512):
        try:
            with torch.inference_mode():
                ext(torch.zeros(1, L))
            short_error[str(L)] = "no error"
        except Exception as e:  # noqa: BLE001
            short_error[str(L)] = type(e).__name__

    manifest = {
        "generated_by": "oracle/gen_golden.py",
        "reference": "AndyWu0719/Audio-CALM preprocess/core.py (MelExtractor, process_audio_chunk), unmodified",
        "torch": torch.__version__, "torchaudio": torchaudio.__version__, "numpy": np.__version__,
        "threads": 1,
        "cases": {k: list(v.shape) for k, v in cases.items()},
        "stats_three_files": stats,
        "frame_table": frame_table,
        "short_input_error": short_error,
        "table_checks": {"window_sum": float(window.sum()), "fb_sum": float(fb.sum()), "fb_max": float(fb.max()),
                         "fb_nonzeros": int((fb != 0).sum())},
    }
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print(json.dumps(manifest["stats_three_files"], indent=1))
    print("wrote", OUT, sum(os.path.getsize(os.path.join(OUT, p)) for p in os.listdir(OUT)), "bytes")


if __name__ == "__main__":
    main()
