"""Mint golden vectors for the Whisper-style preset from the unmodified third-party implementation the reference uses
(transformers.WhisperFeatureExtractor, reached through eval/eval_calm.py:548-552) and from the reference's models/mel_filters.npz.

Run in the build container (needs transformers and /root/reference):  python oracle/gen_golden_whisper.py
Writes tests/golden/whisper_cases.npz (inputs are regenerated from seeds by oracle.logmel_oracle.hash_noise / synth_clip).
"""
import os
import sys

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from oracle import logmel_oracle as o  # noqa: E402


def main() -> None:
    from transformers import WhisperFeatureExtractor
    fe = WhisperFeatureExtractor()
    assert (fe.n_fft, fe.hop_length, fe.feature_size, fe.n_samples) == (400, 160, 80, 480000)
    out = {}
    npz = np.load("/root/reference/models/mel_filters.npz")
    bank = npz["mel_80"].astype(np.float32)                    # (80, 201)
    assert np.abs(bank.T - fe.mel_filters).max() < 1e-8
    # the bank is sparse (<= 2 non-zeros per bin): keep it as (row, col, value) triplets
    r, c = np.nonzero(bank)
    out["bank_rows"], out["bank_cols"], out["bank_vals"] = r.astype(np.int16), c.astype(np.int16), bank[r, c]
    # full 30 s pipeline (pad to 480000), sampled: every 7th frame of every band plus the exact extrema
    cases = {"noise_3s_s5": o.hash_noise(48000, 5), "synth_10s_s7": o.synth_clip(160000, 7), "synth_30s_s9": o.synth_clip(480000, 9),
             "tone_1s": (0.5 * np.sin(2 * np.pi * 440 * np.arange(16000) / 16000) + 0.25 * np.sin(2 * np.pi * 3000 * np.arange(16000) / 16000 + 1)).astype(np.float32)}
    for name, x in cases.items():
        y = fe(x, sampling_rate=16000, return_tensors="np")["input_features"][0]      # (80, 3000) float32
        assert y.shape == (80, 3000)
        out[f"full_{name}_sub7"] = y[:, ::7].astype(np.float32)
        out[f"full_{name}_minmax"] = np.array([y.min(), y.max(), y.mean()], dtype=np.float64)
    # un-padded clips (padding=False): small complete outputs
    for name, x in {"noise_1s_s1": o.hash_noise(16000, 1), "synth_2s_s3": o.synth_clip(32000, 3), "noise_odd_s4": o.hash_noise(20011, 4)}.items():
        y = fe(x, sampling_rate=16000, return_tensors="np", padding=False, truncation=False)["input_features"][0]
        assert y.shape == (80, x.shape[0] // 160), y.shape
        out[f"raw_{name}"] = y.astype(np.float32)
    path = os.path.join(ROOT, "tests", "golden", "whisper_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
