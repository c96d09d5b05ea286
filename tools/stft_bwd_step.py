"""One launch of the STFT-magnitude backward kernel per resolution at the benchmark size: the command ncu profiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audio_calm_b200 as acb
from audio_calm_b200 import spectral
lib = acb._lib.load()
xs = torch.randn(1024, 80, 256, device="cuda") * 3.0 - 6.0
gx = torch.empty_like(xs)
for rep in range(2):
    for n_fft, hop in spectral.STFT_LOSS_SPECS:
        frames = spectral.stft_frames(256, n_fft, hop)
        g = torch.randn((1024, 80, n_fft // 2 + 1, frames), device="cuda")
        w = spectral._window(n_fft, xs.device)
        lib.acb_stft_mag_backward(xs.data_ptr(), g.data_ptr(), 1024 * 80, 256, n_fft, hop, w.data_ptr(), gx.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
