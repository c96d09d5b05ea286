"""Mint golden vectors for the VAE-side spectral op from the UNMODIFIED reference (run in the build container only:
/root/reference does not exist on the GPU box).

    python oracle/gen_golden_spectral.py            # -> tests/golden/stft_mag_cases.npz

Inputs are log-mel features of the committed golden set (tests/golden/cases.npz, themselves minted from the reference) cropped to
the training crop of 256 frames, plus a deterministic 96-frame case that only fits two of the three resolutions; outputs are
AcousticVAE._stft_mag at the three resolutions of stft_loss and the loss value itself (models/modeling_vae.py:271-305).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(1, ROOT)

from models.modeling_vae import AcousticVAE  # noqa: E402  (the reference's own code)
from oracle import logmel_oracle as o  # noqa: E402


def main():
    torch.set_num_threads(1)
    g = np.load(os.path.join(ROOT, "tests", "golden", "cases.npz"))
    mel = g["pipeline_noise_100001_s3"]                      # [80, 392] log-mel minted from the reference pipeline
    x = np.stack([mel[:16, :256], mel[40:56, 100:356]]).astype(np.float32)          # [2, 16, 256]
    y = (x + 0.05 * o.hash_noise(x.size, 77).reshape(x.shape)).astype(np.float32)
    short = g["pipeline_noise_40000_s2"][None, 8:12, :96].astype(np.float32)        # [1, 4, 96]: n_fft 256 does not fit
    out = {"x": x, "y": y, "short": short}
    vae = AcousticVAE.__new__(AcousticVAE)                    # stft_loss only touches self._stft_mag (a static method)
    with torch.no_grad():
        for n_fft, hop in ((256, 64), (128, 32), (64, 16)):
            out[f"mag_x_{n_fft}"] = AcousticVAE._stft_mag(torch.from_numpy(x), n_fft=n_fft, hop_length=hop).numpy()
            if n_fft <= short.shape[-1]:
                out[f"mag_short_{n_fft}"] = AcousticVAE._stft_mag(torch.from_numpy(short), n_fft=n_fft, hop_length=hop).numpy()
        out["loss_xy"] = np.float32(AcousticVAE.stft_loss(vae, torch.from_numpy(x), torch.from_numpy(y)).item())
        out["loss_short"] = np.float32(AcousticVAE.stft_loss(vae, torch.from_numpy(short), torch.from_numpy(short * 0.5)).item())
    path = os.path.join(ROOT, "tests", "golden", "stft_mag_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()}, "torch", torch.__version__)


if __name__ == "__main__":
    main()
