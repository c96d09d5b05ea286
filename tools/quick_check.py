"""Ad-hoc GPU sanity run used during bring-up (not part of the test suite)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import audio_calm_b200 as acb
from audio_calm_b200.preprocess.core import MelExtractor, process_audio_chunk
from oracle import logmel_oracle as o

g = np.load("tests/golden/cases.npz")
ext = MelExtractor().to("cuda").eval()
def run(x):
    return ext(torch.from_numpy(x).cuda()).cpu().numpy()
for name, x in [("raw_noise_16000_s1", o.hash_noise(16000, 1)), ("raw_synth_8000_s11", o.synth_clip(8000, 11)),
                ("raw_synth_513_s13", o.synth_clip(513, 13)), ("raw_synth_1279_s16", o.synth_clip(1279, 16)),
                ("raw_synth_24001_s12", o.synth_clip(24001, 12)), ("raw_zeros_4000", np.zeros(4000, np.float32))]:
    got = run(x[None])[0]
    ref = g[name]
    print(name, got.shape, ref.shape, "maxabs", float(np.max(np.abs(got - ref))) if got.shape == ref.shape else "SHAPE")
fe = acb.LogMelFrontend("cuda")
x = torch.from_numpy(np.stack([o.synth_clip(480000, 100 + i) for i in range(4)])).cuda()
y = fe.forward(x, affine=(acb.MEL_MEAN_DEFAULT, acb.MEL_STD_DEFAULT))
torch.cuda.synchronize()
ref = o.normalise_global(o.logmel(x[1].cpu().numpy(), fe.window.numpy(), fe.fb.numpy()).astype(np.float32))
print("30s clip", y.shape, float(np.max(np.abs(y[1].cpu().numpy() - ref))))
xb = torch.randn(256, 480000, device="cuda") * 0.1
for _ in range(3): fe.forward(xb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): fe.forward(xb)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
frames = 256 * 1876
print(f"256x30s: {ms:.3f} ms  {frames/ms/1e6:.3f} Gframes/s  {256*30/ms*1e3/3600:.1f} audio-h/s  roofline {frames*1344/ms/1e6/6537.6:.3f}")
