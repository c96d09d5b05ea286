"""CPU tests of the Whisper-style preset: the oracle against golden vectors minted from the unmodified
transformers.WhisperFeatureExtractor (oracle/gen_golden_whisper.py), the product's tables against models/mel_filters.npz
(kept as triplets in the golden file), and the C-ABI host logic of the tensor-core route (no GPU compute)."""
import ctypes
import os

import numpy as np
import pytest

import audio_calm_b200 as acb
from oracle import logmel_oracle as o
from oracle import whisper_oracle as wo

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "whisper_cases.npz")
TOL = 1e-4          # north star: log-mel within 1e-4 absolute in fp32


@pytest.fixture(scope="module")
def gw():
    return np.load(GOLDEN)


def golden_bank(gw):
    bank = np.zeros((80, 201), np.float32)
    bank[gw["bank_rows"], gw["bank_cols"]] = gw["bank_vals"]
    return bank


def test_bank_matches_mel_filters_npz(gw):
    bank = golden_bank(gw)
    assert np.abs(wo.mel_filters_f64().T - bank).max() < 1e-8
    _, fb = acb.whisper_tables()
    assert np.abs(fb.numpy().T - bank).max() < 1e-7
    assert (bank[:, 0] == 0).all() and (bank[:, 200] == 0).all()
    assert ((bank != 0).sum(0) <= 2).all()          # triangular: a bin feeds at most two bands


def test_window_is_symmetric():
    w, _ = acb.whisper_tables()
    w = w.numpy()
    assert np.abs(w[1:200] - w[399:200:-1]).max() < 1e-6 and w[0] == 0.0     # fp32 rounding of torch.hann_window: 3e-7
    assert np.abs(w - wo.hann_window_f64()).max() < 1e-6


@pytest.mark.parametrize("name,x", [("noise_1s_s1", lambda: o.hash_noise(16000, 1)), ("synth_2s_s3", lambda: o.synth_clip(32000, 3)),
                                    ("noise_odd_s4", lambda: o.hash_noise(20011, 4))])
def test_oracle_vs_extractor_unpadded(gw, name, x):
    ref = gw[f"raw_{name}"]
    got = wo.whisper_logmel(x())
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 2e-5


@pytest.mark.parametrize("name,x", [("noise_3s_s5", lambda: o.hash_noise(48000, 5)), ("synth_10s_s7", lambda: o.synth_clip(160000, 7)),
                                    ("synth_30s_s9", lambda: o.synth_clip(480000, 9))])
def test_oracle_vs_extractor_30s(gw, name, x):
    got = wo.whisper_logmel(wo.pad_or_trim(x()))
    assert got.shape == (80, 3000)
    assert np.abs(got[:, ::7] - gw[f"full_{name}_sub7"]).max() < 2e-5
    mn, mx, mean = gw[f"full_{name}_minmax"]
    assert abs(got.min() - mn) < 2e-5 and abs(got.max() - mx) < 2e-5 and abs(got.mean() - mean) < 2e-5


def test_tone_floor(gw):
    x = (0.5 * np.sin(2 * np.pi * 440 * np.arange(16000) / 16000) + 0.25 * np.sin(2 * np.pi * 3000 * np.arange(16000) / 16000 + 1)).astype(np.float32)
    got = wo.whisper_logmel(wo.pad_or_trim(x))
    ref = gw["full_tone_1s_sub7"]
    # the zero-padded tail sits exactly on the dynamic-range floor (max - 8) in both
    assert np.abs(got[:, ::7] - ref)[:, 20:].max() < 1e-6
    assert got.max() - got.min() == pytest.approx(2.0, abs=1e-6)      # 8 / 4


def test_frame_counts():
    assert wo.frames_for_length(480000) == 3000 and wo.frames_for_length(480000, drop_last=False) == 3001
    assert wo.frames_for_length(201) == 1
    with pytest.raises(RuntimeError):
        wo.frames_for_length(200)


def test_abi_host_logic(built_lib):
    lib = acb._lib.load()
    assert lib.acb_dftgemm_frames(480000, 1) == 3000 and lib.acb_dftgemm_frames(480000, 0) == 3001
    assert lib.acb_dftgemm_frames(201, 1) == 1 and lib.acb_dftgemm_frames(200, 1) == -1
    for L in (201, 999, 16000, 20011, 480000):
        assert lib.acb_dftgemm_frames(L, 1) == wo.frames_for_length(L)
    # workspace of the dynamic-range floor: one maximum per clip + one minimum per tile of 128 frames
    assert lib.acb_dftgemm_workspace_ints(480000, 1, 256) == 256 * (1 + 24)
    assert lib.acb_dftgemm_workspace_ints(16000, 1, 3) == 3 * (1 + 1) and lib.acb_dftgemm_workspace_ints(200, 1, 3) == -1
    # create() validates before touching the device: wrong transform size, asymmetric window
    h = ctypes.c_void_p()
    w, fb = acb.whisper_tables()
    rc = lib.acb_dftgemm_create(ctypes.byref(h), 0, 1024, 256, 80, w.data_ptr(), fb.data_ptr(), 1e-10, 1)
    assert rc == -3 and b"n_fft=400" in lib.acb_last_error()
    w2 = w.clone()
    w2[7] += 0.01
    rc = lib.acb_dftgemm_create(ctypes.byref(h), 0, 400, 160, 80, w2.data_ptr(), fb.data_ptr(), 1e-10, 1)
    assert rc == -3 and b"symmetric" in lib.acb_last_error()
    dense = fb.clone()
    dense[50, :] = 1.0                       # every band now spans up to 150 bins: too many banded weights for shared memory
    rc = lib.acb_dftgemm_create(ctypes.byref(h), 0, 400, 160, 80, w.data_ptr(), dense.data_ptr(), 1e-10, 1)
    assert rc == -3 and b"dense" in lib.acb_last_error()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(RuntimeError):
        acb.WhisperLogMel("cpu")
