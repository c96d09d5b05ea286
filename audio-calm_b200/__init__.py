"""audio-calm_b200: B200-native (sm_100a) log-mel front-end and mel statistics pass for Audio-CALM.

The directory name carries a hyphen (it is mandated by the build contract), so the importable name is
``audio_calm_b200``: the repository root holds a small ``audio_calm_b200.py`` loader that registers this
directory as that package.  Layout:

    csrc/acb_kernels.cu      hand-written CUDA kernels + the extern "C" ABI (include/audiocalm_b200.h)
    csrc/acb_dftgemm.cu      tensor-core route: STFT as a split-fp16 DFT-GEMM on tcgen05 (Whisper-style preset)
    whisper.py               WhisperLogMel: WhisperFeatureExtractor-compatible features on that route
    _lib.py                  nvcc build + ctypes binding (fails loudly when the .so is missing)
    tables.py                window / slaney filterbank, bit-identical to the reference's torch tables
    frontend.py              LogMelFrontend: batched + ragged launches, fused peak-norm / affine / moments
    stats.py                 per-bin moments, single all-reduce, finalise like compute_mel_stats.py
    sharding.py              utterance sharding across ranks (length-balanced)
    collate.py               training-feed collation on the device (crop / zero-pad to 256 frames; ragged pad + transpose)
    spectral.py              VAE-side STFT magnitudes over the mel time axis (AcousticVAE._stft_mag / stft_loss forward)
    csrc/acb_spectral.cu     its kernel (shares the in-register FFT-32 of csrc/acb_fft32.cuh)
    preprocess/              drop-in mirrors of the reference's preprocess/{core,compute_mel_stats,process_dataset}.py
"""
from . import _lib, tables  # noqa: F401
from .frontend import (LogMelFrontend, MEL_MEAN_DEFAULT, MEL_STD_DEFAULT, RaggedBatch, frames_for_length,  # noqa: F401
                       pack_clips, padded_frames)
from .stats import MelStats, MelStatsAccumulator, finalize_moments, normalize_per_utterance  # noqa: F401
from . import collate, frontend, stats, sharding  # noqa: F401
from .collate import crop_collate, pad_collate, pad_collate_packed  # noqa: F401
from . import spectral, whisper  # noqa: F401
from .whisper import WhisperLogMel, whisper_tables  # noqa: F401

__all__ = ["LogMelFrontend", "MelStatsAccumulator", "MelStats", "RaggedBatch", "pack_clips", "frames_for_length",
           "padded_frames", "finalize_moments", "normalize_per_utterance", "MEL_MEAN_DEFAULT", "MEL_STD_DEFAULT", "WhisperLogMel", "whisper_tables"]
__version__ = "0.1.0"
