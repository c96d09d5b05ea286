"""A few launches of the tensor-core route at the benchmark size (256 x 30 s): the command ncu profiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import audio_calm_b200 as acb
fe = acb.WhisperLogMel("cuda")
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn((256, 480000), device="cuda", generator=g) * 0.1
out = torch.empty((256, 80, 3000), device="cuda")
for _ in range(4):
    fe.forward(x, out=out)
torch.cuda.synchronize()
