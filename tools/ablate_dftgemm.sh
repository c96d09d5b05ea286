# Development: price each part of the tensor-core kernel's K loop in situ (run on a GPU box).  Each variant is built with
# -DACBG_ABLATE=n (results are wrong for n != 0), timed at 256 x 30 s and traced with tools/whisper_trace.py (K-loop cycles per tile).
#   1 no fence.proxy.async   2 no sample loads   3 no operand stores / conversions   4 no MMAs   5 byte permutes instead of conversions
for cfg in "-DACBG_ABLATE=0" "-DACBG_ABLATE=1" "-DACBG_ABLATE=2" "-DACBG_ABLATE=3" "-DACBG_ABLATE=4" "-DACBG_ABLATE=5" "-DACBG_ABLATE=0 -DACBG_A_HI_TMEM=0"; do
  ACB_NVCC_EXTRA="-DACB_DEV $cfg" python -c "import audio_calm_b200 as acb; acb._lib.build(force=True)" || exit 1
  echo "== $cfg"; timeout 100 python tools/whisper_debug.py 2>&1 | tail -2 | head -1; timeout 100 python tools/whisper_trace.py 2>&1 | tail -41 | sed -n 3,4p | cut -c1-64
done
python -c "import audio_calm_b200 as acb; acb._lib.build(force=True)"
