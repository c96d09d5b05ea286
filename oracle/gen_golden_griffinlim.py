"""Mint golden vectors for the Griffin-Lim vocoder fallback from the reference's own call sequence (build container only).

    python oracle/gen_golden_griffinlim.py            # -> tests/golden/griffinlim_cases.npz

eval/eval_calm.py cannot be imported here (it needs peft / evaluate / speechbrain), so the three torchaudio calls of its ``Vocoder``
(``:184-188`` MelScale(n_mels=80, sample_rate=16000, n_stft=513).fb, torch.linalg.pinv, GriffinLim(n_fft=1024); ``:199-208`` decode) are
made exactly as written there, on a log-mel minted from the reference pipeline.  torchaudio's GriffinLim starts from torch.rand phases:
the generator is seeded and the same draw is stored as ``init_angles`` so that the CUDA path and the oracle can start from it."""
import os
import sys

import numpy as np
import torch
import torchaudio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(1, ROOT)


def main():
    torch.set_num_threads(1)
    g = np.load(os.path.join(ROOT, "tests", "golden", "cases.npz"))
    mel = torch.from_numpy(g["pipeline_noise_40000_s2"][None, :, :32].copy())          # [1, 80, 32] log-mel from the reference pipeline
    mel_fb = torchaudio.transforms.MelScale(n_mels=80, sample_rate=16000, n_stft=513).fb   # eval_calm.py:184-186
    inverse_mel_basis = torch.linalg.pinv(mel_fb)                                          # :187
    energy = torch.exp(mel)                                                                # :200
    mag = torch.sqrt(torch.clamp(torch.matmul(energy.transpose(1, 2), inverse_mel_basis).transpose(1, 2), min=1e-8))   # :201-206
    out = {"mel": mel.numpy(), "mag": mag.numpy()}
    for n_iter in (2, 32):
        gl = torchaudio.transforms.GriffinLim(n_fft=1024, n_iter=n_iter)                   # :188 (n_iter = 32 is the default)
        torch.manual_seed(1234)
        init = torch.rand(mag.size(), dtype=torch.complex64)                               # the draw griffinlim makes first
        torch.manual_seed(1234)
        wav = gl(mag).squeeze(1)                                                           # :208
        out["init_angles"] = init.numpy()                                                  # the same draw for every n_iter (same seed)
        out[f"wave_{n_iter}"] = wav.numpy()
    # one torch.istft / torch.stft round trip on its own
    w = torch.hann_window(1024)
    spec = torch.stft(torch.from_numpy(np.random.default_rng(5).normal(0, 0.1, (1, 6000)).astype(np.float32)), 1024, 512, 1024, w, center=True,
                      pad_mode="reflect", return_complex=True)
    out["rt_spec"] = spec.numpy()
    out["rt_wave"] = torch.istft(spec, 1024, 512, 1024, w).numpy()
    path = os.path.join(ROOT, "tests", "golden", "griffinlim_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()}, "torch", torch.__version__, "torchaudio", torchaudio.__version__)


if __name__ == "__main__":
    main()
