"""Development check of the tensor-core route on a GPU box: localises errors by tile / band range and times the kernel."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import audio_calm_b200 as acb
from oracle import logmel_oracle as o, whisper_oracle as wo

fe = acb.WhisperLogMel("cuda", dyn_range=0.0, affine_mean=None, drop_last_frame=False)
for L in (16000, 70000):
    x = o.hash_noise(L, 3)
    try:
        y = fe.forward(torch.from_numpy(x).cuda()[None], check=True)[0].cpu().numpy()
    except Exception as e:  # noqa: BLE001
        print("L", L, "FAILED:", e); continue
    ref = wo.whisper_logmel(x, fe.window.numpy(), fe.fb.numpy(), drop_last=False, dyn_range=None, affine=False)
    d = np.abs(y - ref)
    print(f"L={L} shape {y.shape} max err {d.max():.3e} nan {np.isnan(y).sum()}")
    T = y.shape[1]
    for t0 in range(0, T, 128):
        blk = d[:, t0:t0 + 128]
        print(f"  tile {t0//128}: max {blk.max():.3e} | bands 0-39 {blk[:40].max():.3e} 40-79 {blk[40:].max():.3e} | frames first8 {blk[:, :8].max():.3e} last8 {blk[:, -8:].max():.3e}")
    if d.max() > 1e-3:
        b, t = np.unravel_index(np.argmax(d), d.shape)
        print("  worst at band", b, "frame", t, "got", y[b, t], "ref", ref[b, t])
        print("  per-band max err:", np.array2string(d.max(1), precision=2, max_line_width=200))
fe2 = acb.WhisperLogMel("cuda")
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn((256, 480000), device="cuda", generator=g) * 0.1
out = torch.empty((256, 80, 3000), device="cuda")
for _ in range(3): fe2.forward(x, out=out)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(10): fe2.forward(x, out=out)
ev[1].record(); torch.cuda.synchronize()
ms = ev[0].elapsed_time(ev[1]) / 10
fr = 256 * 3000
print(f"256 x 30 s: {ms:.3f} ms/step (both kernels) -> {fr/ms/1e6:.3f} G frames/s, {256*30/ms*1000/3600:.0f} audio-h/s, algorithmic {fr*960/ms/1e6:.0f} GB/s")
ref = wo.whisper_logmel(x[5].cpu().numpy(), fe2.window.numpy(), fe2.fb.numpy())
print("bench-size clip 5 max err", np.abs(out[5].cpu().numpy() - ref).max())
