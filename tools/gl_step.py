"""A few Griffin-Lim iterations at 64 clips x 626 frames: the command ncu profiles (acb_istft, acb_stft_complex)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio_calm_b200 import spectral
x = torch.rand(64, 513, 626, device="cuda") + 0.01
spectral.griffin_lim(x, n_iter=3)
torch.cuda.synchronize()
