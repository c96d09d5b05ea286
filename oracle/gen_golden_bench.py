"""Mint golden vectors at the BENCHMARK's own size and distribution from the unmodified reference (build container only).

    python oracle/gen_golden_bench.py            # -> tests/golden/bench_cases.npz

Three 30 s clips of ``logmel_oracle.bench_clip`` (the RNG-independent twin of ``bench.synth_batch``) go through the reference's
``MelExtractor`` (preprocess/core.py:50-61) and the VAE's scalar normalisation (models/modeling_vae.py:317-319) -- configs[1] of
BASELINE.json.  To keep the fixture small every 7th frame is stored, next to the minimum, maximum and mean of the full result."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(1, ROOT)

from preprocess.core import MelExtractor  # noqa: E402  (the reference itself)
from oracle.logmel_oracle import bench_clip  # noqa: E402


def main():
    torch.set_num_threads(1)
    ext = MelExtractor().eval()
    out = {}
    with torch.inference_mode():
        for seed in (101, 202, 303):
            x = bench_clip(480000, seed)
            mel = (ext(torch.from_numpy(x)[None]) - (-6.589515)) / 3.860679
            y = mel[0].contiguous().numpy()
            assert y.shape == (80, 1876)
            out[f"norm_bench_30s_s{seed}_sub7"] = y[:, ::7].copy()
            out[f"norm_bench_30s_s{seed}_stats"] = np.array([y.min(), y.max(), y.astype(np.float64).mean()], dtype=np.float64)
    path = os.path.join(ROOT, "tests", "golden", "bench_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()}, "torch", torch.__version__)


if __name__ == "__main__":
    main()
