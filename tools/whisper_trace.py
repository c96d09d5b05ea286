"""Development (library built with ACB_NVCC_EXTRA=-DACB_DEV): per-phase clock stamps of CTA 0 of the tensor-core kernel (worker warp 0, MMA issuer, loader) for the first tiles."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import audio_calm_b200 as acb
if os.environ.get("ACB_LIB"):
    acb._lib.LIB_PATH = os.environ["ACB_LIB"]
fe = acb.WhisperLogMel("cuda")
lib = acb._lib.load()
lib.acb_dftgemm_set_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn((256, 480000), device="cuda", generator=g) * 0.1
out = torch.empty((256, 80, 3000), device="cuda")
for _ in range(3): fe.forward(x, out=out)
tr = torch.zeros(3 * 48 * 16, dtype=torch.int64, device="cuda")
lib.acb_dftgemm_set_trace(fe._handle, tr.data_ptr())
fe.forward(x, out=out, check=True)
lib.acb_dftgemm_set_trace(fe._handle, None)
t = tr.cpu().numpy().reshape(3, 48, 16)
t0 = t[0, 0, 0]
names_w = ["smp", "A0", "A1", "A2", "A3", "A4", "A5", "A6", "tile_done", "pow_sync", "mel_end", "end_sync"]
for ti in range(1, 6):
    print(f"tile {ti}: worker  " + " ".join(f"{n}={t[0, ti, i] - t0}" for i, n in enumerate(names_w)))
    print(f"        issuer  " + " ".join(f"a{k}={t[1, ti, 2*k] - t0} b{k}={t[1, ti, 2*k+1] - t0}" for k in range(7)))
    print(f"        loader  " + " ".join(f"B{k}={t[2, ti, k] - t0}" for k in range(7)) + f" sfree={t[2, ti, 8] - t0}")

print("per tile: period | K loop (A6 - smp) | accumulators (tile_done - A6) | power | mel | gap to next smp")
for ti in range(0, 41):
    w = t[0, ti]
    nxt = t[0, ti + 1, 0] if t[0, ti + 1, 0] else 0
    print(f"{ti:2d}: {nxt - w[0] if nxt else 0:6d} | {w[7] - w[0]:6d} | {w[8] - w[7]:6d} | {w[9] - w[8]:5d} | {w[10] - w[9]:5d} | {nxt - w[11] if nxt else 0:6d}   B-ready minus A-ready per step: " + " ".join(str(t[1, ti, 2*k+1] - t[1, ti, 2*k]) for k in range(7)))
