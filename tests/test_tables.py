"""Window / filterbank tables must be bit-identical to the buffers the reference's MelExtractor holds."""
import numpy as np
import pytest
import torch

import audio_calm_b200 as acb
from audio_calm_b200 import tables
from oracle import logmel_oracle as o


def test_tables_bitwise_vs_reference(golden_tables, manifest):
    w, fb = tables.calm_tables()
    assert np.array_equal(w.numpy(), golden_tables["window"])
    assert np.array_equal(fb.numpy(), golden_tables["fb"])
    chk = manifest["table_checks"]
    assert abs(float(w.numpy().sum(dtype=np.float64)) - 512.0) < 1e-4 and abs(chk["window_sum"] - 512.0) < 1e-4
    assert int((fb != 0).sum()) == chk["fb_nonzeros"] == 1001
    assert abs(float(fb.sum()) - chk["fb_sum"]) < 1e-6


def test_tables_vs_torchaudio_if_present():
    ta = pytest.importorskip("torchaudio")
    for sr, n_fft, n_mels, f_max in ((16000, 1024, 80, 8000.0), (16000, 512, 40, 8000.0), (22050, 1024, 80, 8000.0)):
        ref = ta.functional.melscale_fbanks(n_fft // 2 + 1, 0.0, f_max, n_mels, sr, norm="slaney", mel_scale="slaney")
        assert torch.equal(ref, tables.slaney_fbanks(n_fft // 2 + 1, 0.0, f_max, n_mels, sr))


def test_banded_form_round_trip():
    _, fb = tables.calm_tables()
    b = tables.band_filterbank(fb)
    assert b.n_mels == 80 and b.weights.size == 1001
    assert int(b.length.min()) == 4 and int(b.length.max()) == 37
    assert list(b.start[:4]) == [1, 3, 5, 8] and list(b.length[:4]) == [4, 5, 5, 4]
    assert int(b.start[-1]) == 475 and int(b.length[-1]) == 37
    assert np.array_equal(b.dense(), fb.numpy())
    assert np.all(fb.numpy()[0] == 0) and np.all(fb.numpy()[512] == 0)      # DC and Nyquist never contribute


def test_fp64_tables_close():
    w, fb = tables.calm_tables()
    assert np.max(np.abs(o.hann_window_f64() - w.numpy())) < 2e-7
    f64 = o.slaney_fbanks_f64()
    assert np.max(np.abs(f64 - fb.numpy())) < 1e-6


def test_module_buffers_and_state_dict():
    from audio_calm_b200.preprocess.core import MelExtractor
    m = MelExtractor()
    sd = m.state_dict()
    assert list(sd.keys()) == ["mel_transform.spectrogram.window", "mel_transform.mel_scale.fb"]
    assert sd["mel_transform.mel_scale.fb"].shape == (513, 80) and sd["mel_transform.spectrogram.window"].shape == (1024,)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4000))                      # CPU tensor: no CPU fallback
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4000, dtype=torch.float64))
