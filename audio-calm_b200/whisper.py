"""Whisper-style log-mel features on the tensor-core route (DFT-as-GEMM on tcgen05; csrc/acb_dftgemm.cu).

BASELINE.json's north star words the front-end as "Whisper-style": n_fft 400, hop 160, the 80-band bank of
``models/mel_filters.npz``, ``log10(clamp(., 1e-10))``, the ``max - 8`` dynamic-range floor and ``(x + 4) / 4``.  In the reference
that computation runs inside the Hugging Face ASR pipeline that scores generated speech (eval/eval_calm.py:548-552 ->
``transformers.WhisperFeatureExtractor``); ``WhisperLogMel`` reproduces ``WhisperFeatureExtractor.__call__`` for fp32 waveforms
on a CUDA device.  The reference's own extractor (``MelExtractor``, n_fft 1024) is ``LogMelFrontend``.

There is no CPU fallback: construction needs a CUDA device and the built C-ABI library.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Union

import torch

from . import _lib
from ._lib import ACB_LOG_10, ACB_LOG_NATURAL, DftGemmArgs
from .tables import hann_window, slaney_fbanks

WHISPER_N_FFT = 400
WHISPER_HOP = 160
WHISPER_N_MELS = 80
WHISPER_CHUNK_SAMPLES = 480000      # 30 s at 16 kHz: WhisperFeatureExtractor pads / trims every clip to this
MAX_ABS_SAMPLE = 2.0                # fp16 operand range of the split-precision GEMM after its 2^12 pre-scale (audio is in [-1, 1])


def whisper_tables(n_mels: int = WHISPER_N_MELS, sample_rate: int = 16000):
    """(window ``[400]``, filterbank ``[201, n_mels]``) fp32: periodic Hann and the slaney-scale, slaney-normalised
    triangular bank over 0-8000 Hz -- the published construction of ``models/mel_filters.npz`` (``mel_80``, stored
    transposed there; tests compare the two to 1e-7)."""
    return hann_window(WHISPER_N_FFT), slaney_fbanks(WHISPER_N_FFT // 2 + 1, 0.0, sample_rate / 2.0, n_mels, sample_rate)


_DEFAULT = object()


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class WhisperLogMel:
    """Device tables + launches of the tcgen05 DFT-GEMM front-end for one device.

    ``forward(wav[B, L])`` -> ``[B, n_mels, L // 160]`` fp32, the ``input_features`` of ``WhisperFeatureExtractor`` for clips
    already padded / trimmed to ``L`` samples; ``extract(clips)`` pads or trims to 30 s first, like ``__call__`` does.
    """

    def __init__(self, device: Union[str, torch.device, int] = "cuda", n_mels: int = WHISPER_N_MELS, clamp_min: float = 1e-10,
                 log: str = "log10", dyn_range: float = 8.0, affine_mean: Optional[float] = -4.0, affine_std: float = 4.0,
                 drop_last_frame: bool = True, window: Optional[torch.Tensor] = None, fb: Optional[torch.Tensor] = None):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("WhisperLogMel runs on a CUDA device (sm_100a, tcgen05); there is no CPU fallback")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self.n_fft, self.hop, self.n_mels = WHISPER_N_FFT, WHISPER_HOP, n_mels
        self.clamp_min, self.dyn_range = float(clamp_min), float(dyn_range)
        self.affine_mean, self.affine_std = affine_mean, float(affine_std)
        self.drop_last_frame = bool(drop_last_frame)
        self.log_kind = {"ln": ACB_LOG_NATURAL, "log10": ACB_LOG_10}[log]
        w0, f0 = whisper_tables(n_mels)
        self.window = (w0 if window is None else window).detach().to("cpu", torch.float32).contiguous()
        self.fb = (f0 if fb is None else fb).detach().to("cpu", torch.float32).contiguous()
        if tuple(self.window.shape) != (self.n_fft,) or tuple(self.fb.shape) != (self.n_fft // 2 + 1, n_mels):
            raise ValueError("window must be [400] and fb [201, n_mels]")
        self._lib = _lib.load()
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            _lib.check(self._lib.acb_dftgemm_create(ctypes.byref(handle), device.index, self.n_fft, self.hop, n_mels,
                                                    self.window.data_ptr(), self.fb.data_ptr(), self.clamp_min, self.log_kind),
                       "acb_dftgemm_create")
        self._handle = handle
        self._clip_max: Optional[torch.Tensor] = None
        self._moments_ws: Optional[torch.Tensor] = None
        self.launches = 0

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                self._lib.acb_dftgemm_destroy(h)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass
            self._handle = None

    def frames_for_length(self, length: int) -> int:
        t = int(self._lib.acb_dftgemm_frames(int(length), int(self.drop_last_frame)))
        if t < 0:
            raise RuntimeError(f"Argument #4: Padding size should be less than the corresponding input dimension, but got: "
                               f"padding ({self.n_fft // 2}, {self.n_fft // 2}) at dimension 2 of input of length {length}")
        return t

    def peak_abs(self, wav: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Per-clip ``max|x|`` (``acb_peak_abs``; over each clip's own ``lengths[i]`` samples when given)."""
        n_clips = int(wav.shape[0])
        peak = torch.empty(n_clips, dtype=torch.float32, device=self.device)
        offs = lens = None
        if lengths is not None:
            lens = lengths.to(self.device, torch.int64).contiguous()
            offs = (torch.arange(n_clips, dtype=torch.int64, device=self.device) * int(wav.stride(0))).contiguous()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.acb_peak_abs(wav.data_ptr(), None if offs is None else offs.data_ptr(), None if lens is None else lens.data_ptr(),
                                              int(wav.stride(0)), int(wav.shape[1]), n_clips, peak.data_ptr(), _stream_ptr(self.device)), "acb_peak_abs")
        self.launches += 1
        return peak

    def forward(self, wav: torch.Tensor, out: Optional[torch.Tensor] = None, check: bool = False,
                out_dtype: torch.dtype = torch.float32, *, lengths: Optional[torch.Tensor] = None, peak: Union[None, bool, torch.Tensor] = None,
                peak_norm: bool = False, affine=_DEFAULT, moments=None, fill_value: float = 0.0) -> torch.Tensor:
        """``wav`` fp32 CUDA ``[B, L]`` (rows may be strided) -> ``[B, n_mels, frames]`` fp32 (or bf16: ``out_dtype`` / the dtype of ``out``).

        ``lengths[B]``: ragged batch -- clip ``i`` holds ``lengths[i]`` samples of its (zero-padded) row; reflection happens at each
        clip's own end and frames beyond its own count are ``fill_value``.
        ``peak``: per-clip ``max|x|`` tensor, or ``True`` to compute it here (one extra pass).  With it every clip is pre-scaled by an
        exact power of two, so ANY amplitude is safe (un-normalised PCM-scale floats included); without it the caller vouches for
        ``|x| <= 2`` and a violation is reported by ``check`` / ``extract`` / ``forward_host`` as an error instead of garbage.
        ``peak_norm``: fused ``process_audio_chunk`` gain ``0.95 / (peak + 1e-8)`` (needs ``peak``).
        ``affine``: ``(mean, std)`` scalars, per-bin tensors (e.g. ``MelStats.affine()``), ``None``; default = the constructor's pair.
        ``moments``: a ``MelStatsAccumulator`` updated with the un-normalised values -- fused when ``dyn_range == 0``; with the
        per-clip floor the features must be stored un-normalised (``affine=None``) and are reduced by ``acb_moments_accumulate``.
        """
        if not wav.is_cuda or wav.device != self.device:
            raise RuntimeError(f"expected a CUDA tensor on {self.device}, got {wav.device} (no CPU fallback)")
        if wav.dtype != torch.float32:
            raise RuntimeError(f"expected float32 samples, got {wav.dtype}")
        if wav.dim() != 2 or wav.stride(1) != 1:
            raise ValueError("wav must be [B, L] with unit stride along time")
        n_clips, length = int(wav.shape[0]), int(wav.shape[1])
        frames = self.frames_for_length(length)
        if out is None:
            out = torch.empty((n_clips, self.n_mels, frames), dtype=out_dtype, device=self.device)
        if out.dtype not in (torch.float32, torch.bfloat16) or out.device != self.device or tuple(out.shape[:2]) != (n_clips, self.n_mels) \
                or out.shape[2] < frames or out.stride(2) != 1 or out.stride(1) != out.shape[2]:
            raise ValueError("out must be a float32 / bfloat16 [B, n_mels, >= frames] tensor with contiguous rows on the same device")
        if n_clips == 0 or frames == 0:
            return out[:, :, :frames]
        keep = []
        a = DftGemmArgs()
        a.wav = wav.data_ptr()
        a.clip_stride = int(wav.stride(0)) if n_clips > 1 else length
        a.length = length
        a.n_clips = n_clips
        a.drop_last_frame = int(self.drop_last_frame)
        a.out = out.data_ptr()
        a.out_dtype = _lib.ACB_BF16 if out.dtype == torch.bfloat16 else _lib.ACB_F32
        a.out_clip_stride = int(out.stride(0)) if n_clips > 1 else self.n_mels * int(out.shape[2])
        a.frame_capacity = int(out.shape[2])
        a.dyn_range = self.dyn_range
        a.fill_value = float(fill_value)
        frames_per_clip = None
        if lengths is not None:
            lens_host = lengths.detach().to("cpu", torch.int64)
            if lens_host.numel() != n_clips or int(lens_host.max()) > length:
                raise ValueError("lengths must hold one sample count <= L per clip")
            if int(lens_host.min()) <= self.n_fft // 2:
                self.frames_for_length(int(lens_host.min()))      # raises like the reference's reflect padding
            lens_dev = lens_host.to(self.device, non_blocking=True).contiguous()
            keep.append(lens_dev)
            a.clip_length = lens_dev.data_ptr()
            frames_per_clip = lens_host // self.hop + (0 if self.drop_last_frame else 1)
        if peak is True:
            peak = self.peak_abs(wav, lengths)
        if peak is not None and peak is not False:
            peak = peak.to(self.device, torch.float32).contiguous()
            keep.append(peak)
            a.clip_peak = peak.data_ptr()
        if peak_norm:
            if not a.clip_peak:
                raise ValueError("peak_norm needs peak (a tensor from peak_abs, or True)")
            a.peak_norm = 1
        if affine is _DEFAULT:
            affine = None if self.affine_mean is None else (self.affine_mean, self.affine_std)
        if affine is None:
            a.affine = 0
        elif isinstance(affine[0], torch.Tensor):
            mean = affine[0].detach().to(self.device, torch.float32).contiguous()
            std = affine[1].detach().to(self.device, torch.float32).contiguous()
            if mean.numel() != self.n_mels or std.numel() != self.n_mels:
                raise ValueError("per-bin affine needs n_mels means and stds")
            keep += [mean, std]
            a.affine, a.bin_mean, a.bin_std = 2, mean.data_ptr(), std.data_ptr()
        else:
            a.affine, a.affine_mean, a.affine_std = 1, float(affine[0]), float(affine[1])
        fused_moments = moments is not None and not self.dyn_range > 0
        if moments is not None and not fused_moments and a.affine != 0:
            raise ValueError("moments of the floored features need them stored un-normalised: pass affine=None")
        if fused_moments:
            if self._moments_ws is None:
                n = int(self._lib.acb_dftgemm_moments_workspace_bytes(self._handle))
                self._moments_ws = torch.empty(n // 8, dtype=torch.float64, device=self.device)
            a.moments = moments.moments.data_ptr()
            a.moments_workspace = self._moments_ws.data_ptr()
        if self.dyn_range > 0:
            need = int(self._lib.acb_dftgemm_workspace_ints(length, int(self.drop_last_frame), n_clips))
            if self._clip_max is None or self._clip_max.numel() < need:
                self._clip_max = torch.empty(need, dtype=torch.int32, device=self.device)
            a.clip_max = self._clip_max.data_ptr()
        stream = _stream_ptr(self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.acb_dftgemm_forward(self._handle, ctypes.byref(a), stream), "acb_dftgemm_forward")
            self.launches += 1 + int(self.dyn_range > 0) + int(fused_moments)
            if check:
                _lib.check(self._lib.acb_dftgemm_check(self._handle, stream), "acb_dftgemm_check")
        res = out[:, :, :frames]
        if moments is not None:
            if fused_moments:
                moments.frames += int(frames_per_clip.sum()) if frames_per_clip is not None else n_clips * frames
            else:
                moments.update(res if res.is_contiguous() else res.contiguous(),
                               None if frames_per_clip is None else frames_per_clip)
        del keep
        return res

    def check(self) -> None:
        """Synchronise and raise if a launch since the last check flagged a barrier timeout or non-finite features
        (``acb_dftgemm_check``)."""
        with torch.cuda.device(self.device):
            _lib.check(self._lib.acb_dftgemm_check(self._handle, _stream_ptr(self.device)), "acb_dftgemm_check")

    __call__ = forward

    def forward_host(self, wav_host: torch.Tensor, out_host: Optional[torch.Tensor] = None, n_chunks: int = 8) -> torch.Tensor:
        """Host buffers in and out (pinned memory for overlap): the batch is cut into ``n_chunks`` groups of clips; the H2D copy of
        group i+1, the kernels of group i and the D2H copy of group i-1 run on three streams.  Returns ``out_host``
        (``[B, n_mels, frames]`` fp32) after synchronising."""
        if wav_host.is_cuda or wav_host.dtype != torch.float32 or wav_host.dim() != 2:
            raise ValueError("wav_host must be a float32 host tensor [B, L]")
        n_clips, length = int(wav_host.shape[0]), int(wav_host.shape[1])
        frames = self.frames_for_length(length)
        if out_host is None:
            out_host = torch.empty((n_clips, self.n_mels, frames), dtype=torch.float32, pin_memory=True)
        n_chunks = max(1, min(int(n_chunks), n_clips))
        bounds = [n_clips * i // n_chunks for i in range(n_chunks + 1)]
        st = getattr(self, "_host_state", None)
        if st is None or st[0].shape != (n_clips, length):
            st = (torch.empty((n_clips, length), dtype=torch.float32, device=self.device),
                  torch.empty((n_clips, self.n_mels, frames), dtype=torch.float32, device=self.device),
                  torch.cuda.Stream(self.device), torch.cuda.Stream(self.device), torch.cuda.Stream(self.device))
            self._host_state = st
        dev_in, dev_out, s_in, s_run, s_out = st
        cur = torch.cuda.current_stream(self.device)
        for s_ in (s_in, s_run, s_out):
            s_.wait_stream(cur)
        for i in range(n_chunks):
            a, b = bounds[i], bounds[i + 1]
            if a == b:
                continue
            with torch.cuda.stream(s_in):
                dev_in[a:b].copy_(wav_host[a:b], non_blocking=True)
                e_in = torch.cuda.Event()
                e_in.record(s_in)
            with torch.cuda.stream(s_run):
                s_run.wait_event(e_in)
                self.forward(dev_in[a:b], out=dev_out[a:b])
                e_run = torch.cuda.Event()
                e_run.record(s_run)
            with torch.cuda.stream(s_out):
                s_out.wait_event(e_run)
                out_host[a:b].copy_(dev_out[a:b], non_blocking=True)
        for s_ in (s_in, s_run, s_out):
            cur.wait_stream(s_)
        self.check()                                   # synchronises; a barrier timeout or an out-of-range sample raises here
        return out_host

    def extract(self, clips: Sequence[torch.Tensor], n_samples: int = WHISPER_CHUNK_SAMPLES, check: bool = True) -> torch.Tensor:
        """``WhisperFeatureExtractor.__call__`` semantics: every clip is zero-padded or trimmed to ``n_samples`` (30 s) before the
        STFT, so the result is ``[len(clips), n_mels, n_samples // 160]``.  Like the extractor it takes ANY float amplitude: the
        per-clip peak is measured and every clip pre-scaled by a power of two (exact)."""
        batch = torch.zeros((len(clips), n_samples), dtype=torch.float32, device=self.device)
        for i, c in enumerate(clips):
            c = c.reshape(-1)[:n_samples]
            batch[i, : c.numel()] = c.to(self.device, torch.float32)
        return self.forward(batch, check=check, peak=True)

    def forward_ragged(self, clips: Sequence[torch.Tensor], **kw):
        """Variable-length clips -> (``[B, n_mels, frames of the longest]``, ``frames[B]``): every clip keeps its own frame count and
        its own reflected end (unlike :meth:`extract`, which pads to 30 s first); the rest of a shorter row is ``fill_value``."""
        lens = torch.tensor([int(c.numel()) for c in clips], dtype=torch.int64)
        L = (int(lens.max()) + 31) // 32 * 32                          # whole 128-byte rows: the tensor-copy path
        batch = torch.zeros((len(clips), L), dtype=torch.float32, device=self.device)
        for i, c in enumerate(clips):
            batch[i, : c.numel()] = c.reshape(-1).to(self.device, torch.float32)
        out = self.forward(batch, lengths=lens, **kw)
        frames = lens // self.hop + (0 if self.drop_last_frame else 1)
        return out[:, :, : int(frames.max())], frames
