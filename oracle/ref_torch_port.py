"""CPU port of the reference's MelExtractor op sequence, for TIMING the reference's CPU path on the GPU box
(`bench.py` ``cpu_baseline`` and ``--impl reference``) and as a second checker next to the numpy oracle.

TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product package.

The reference tree (/root/reference) does not exist on the GPU box, so its ``preprocess/core.py`` cannot be
imported there.  That file is a thin wrapper: ``torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=1024,
hop_length=256, n_mels=80, power=2.0, normalized=False, f_min=0, f_max=8000, norm="slaney", mel_scale="slaney")``
followed by ``torch.log(torch.clamp(mel, min=1e-5))`` (preprocess/core.py:37-48, 55, 60).  This module restates
exactly those calls (torchaudio is part of the image); if torchaudio is missing it falls back to ``torch.stft`` plus
the filterbank matmul, i.e. the same ATen kernels torchaudio dispatches to.  In the build container the port is
checked bit-for-bit against the imported reference by ``oracle/gen_golden.py`` outputs (tests/test_ref_port.py).
"""
from __future__ import annotations

import torch

try:
    import torchaudio
    _HAVE_TORCHAUDIO = True
except Exception:  # noqa: BLE001
    torchaudio = None
    _HAVE_TORCHAUDIO = False


class RefMelExtractor(torch.nn.Module):
    """preprocess/core.py:23-61 restated."""

    def __init__(self, sample_rate=16000, n_fft=1024, hop_length=256, n_mels=80):
        super().__init__()
        self.n_fft, self.hop_length = n_fft, hop_length
        if _HAVE_TORCHAUDIO:
            self.mel_transform = torchaudio.transforms.MelSpectrogram(
                sample_rate=sample_rate, n_fft=n_fft, hop_length=hop_length, n_mels=n_mels, power=2.0, normalized=False,
                f_min=0, f_max=8000, norm="slaney", mel_scale="slaney")
        else:
            from oracle.logmel_oracle import slaney_fbanks_f64
            self.mel_transform = None
            self.register_buffer("window", torch.hann_window(n_fft))
            self.register_buffer("fb", torch.from_numpy(slaney_fbanks_f64(n_fft // 2 + 1, 0.0, 8000.0, n_mels, sample_rate)).float())

    def forward(self, wav):
        if self.mel_transform is not None:
            mel = self.mel_transform(wav)
        else:
            shape = wav.shape
            spec = torch.stft(wav.reshape(-1, shape[-1]), self.n_fft, self.hop_length, self.n_fft, self.window, center=True,
                              pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
            p = spec.abs().pow(2.0)
            mel = torch.matmul(p.transpose(-1, -2), self.fb).transpose(-1, -2)
            mel = mel.reshape(shape[:-1] + mel.shape[-2:])
        return torch.log(torch.clamp(mel, min=1e-5))


def process_audio_chunk(wav, target_sr=16000):
    """preprocess/core.py:93-112 restated."""
    if wav.shape[0] > 1:
        wav = torch.mean(wav, dim=0, keepdim=True)
    peak = torch.max(torch.abs(wav))
    if peak > 0:
        wav = wav / (peak + 1e-8) * 0.95
    return wav


def normalise(mel, mel_mean=-6.589515, mel_std=3.860679):
    """models/modeling_vae.py:317-319."""
    return (mel - mel_mean) / mel_std
