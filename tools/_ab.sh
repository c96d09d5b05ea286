for cfg in "-DACBG_WORKER_WARPS=8 -DACBG_BAND_COST=8" "-DACBG_WORKER_WARPS=8 -DACBG_BAND_COST=20" "-DACBG_WORKER_WARPS=16 -DACBG_BAND_COST=20" "-DACBG_WORKER_WARPS=16 -DACBG_BAND_COST=32"; do
  ACB_NVCC_EXTRA="$cfg" python -c "import audio_calm_b200 as acb; acb._lib.build(force=True)" || exit 1
  echo "== $cfg"; timeout 100 python tools/whisper_debug.py 2>&1 | tail -2 | head -1; timeout 100 python tools/whisper_trace.py 2>&1 | grep -A0 "tile 2: worker"
done
