"""In-situ cost of each part of the fused log-mel kernel (development tool, not part of the product path).

    python tools/ablate.py build            # here (CPU box): nvcc -DACB_ABLATE=n  ->  tools/_ablate/lib<n>.so
    python tools/ablate.py run [n ...]      # on the GPU box: time config 2 (256 x 30 s) with every variant

A variant removes ONE piece of the kernel (its output is wrong on purpose); the time it saves is what that piece costs
inside the real instruction mix -- the evidence DESIGN.md quotes for where the cycles go.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tools", "_ablate")
NAMES = {0: "baseline", 4: "no mel inner loop", 7: "no phase 1 (FFT) at all", 8: "no phase 2 (mel/log/store) at all"}


def build():
    import audio_calm_b200 as acb
    os.makedirs(OUT, exist_ok=True)
    for n in NAMES:
        dst = os.path.join(OUT, f"lib{n}.so")
        cmd = [acb._lib._nvcc()] + acb._lib.NVCC_FLAGS + ["-DACB_DEV", f"-DACB_ABLATE={n}", "-I", acb._lib.INCLUDE, "-o", dst,
                                                         os.path.join(acb._lib.CSRC, "acb_kernels.cu")]
        subprocess.run(cmd, check=True)
        print("built", dst)


def run_one(n: int):
    import torch
    import audio_calm_b200 as acb
    acb._lib.LIB_PATH = os.path.join(OUT, f"lib{n}.so")
    fe = acb.LogMelFrontend("cuda")
    x = torch.randn(256, 480000, device="cuda") * 0.1
    out = torch.empty((256, 80, 1876), device="cuda")
    for _ in range(3):
        fe.forward(x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fe.forward(x, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"ablate {n}: {ms:.4f} ms  {256 * 1876 / ms / 1e6:.3f} Gframes/s   [{NAMES[n]}]", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    elif sys.argv[1] == "one":
        run_one(int(sys.argv[2]))
    else:
        ids = [int(a) for a in sys.argv[2:]] or sorted(NAMES)
        for n in ids:   # one process per variant: the library is loaded once per process
            subprocess.run([sys.executable, os.path.abspath(__file__), "one", str(n)], check=False)
