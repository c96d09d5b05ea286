"""32 Griffin-Lim iterations (n_fft 1024, hop 512, momentum 0.99) over a few batch shapes: spectral.griffin_lim (acb_istft +
acb_stft_complex with the fused phase update) next to torchaudio.transforms.GriffinLim (cuFFT) on the same GPU.  One JSON line each."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torchaudio
import audio_calm_b200 as acb
if os.environ.get("ACB_LIB"):
    acb._lib.LIB_PATH = os.environ["ACB_LIB"]
from audio_calm_b200 import spectral


def wall_ms(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t) * 1e3)
    return best


gl = torchaudio.transforms.GriffinLim(n_fft=1024).cuda()
for B, T in ((1, 626), (8, 626), (64, 626), (8, 3126)):
    x = torch.rand(B, 513, T, device="cuda") + 0.01
    ours, ref = wall_ms(lambda: spectral.griffin_lim(x)), wall_ms(lambda: gl(x))
    print(json.dumps({"tool": "griffinlim_sweep", "clips": B, "frames": T, "seconds_per_clip": round(512 * (T - 1) / 16000, 1), "ours_ms": ours,
                      "torchaudio_ms": ref, "speedup": ref / ours}), flush=True)
