"""Host-side logic and the C-ABI surface, checked without a GPU (only pure-host entry points are called)."""
import ctypes
import os
import re

import numpy as np
import pytest

import audio_calm_b200 as acb
from audio_calm_b200 import _lib, sharding, stats
from oracle import logmel_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "audiocalm_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(acb_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    lib = ctypes.CDLL(built_lib)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert _lib.load().acb_abi_version() == 2


def test_args_struct_matches_header():
    header = open(os.path.join(ROOT, "include", "audiocalm_b200.h")).read()
    body = header[header.index("typedef struct acb_logmel_args {"):header.index("} acb_logmel_args;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\b([a-z_0-9]+)\s*;", body)
    assert names == [f[0] for f in _lib.LogmelArgs._fields_]


def test_frame_arithmetic_c_abi(built_lib, manifest):
    lib = _lib.load()
    for L, T in manifest["frame_table"].items():
        assert lib.acb_frames_for_length(int(L), 1024, 256) == T == acb.frames_for_length(int(L))
    for L in (0, 1, 256, 512):
        assert lib.acb_frames_for_length(L, 1024, 256) == -1
        with pytest.raises(RuntimeError):
            acb.frames_for_length(L)
    for T in range(1, 40):
        assert lib.acb_padded_frames(T, 4) == o.padded_frames(T, 4) == acb.padded_frames(T, 4)
        assert lib.acb_padded_frames(T, 1) == T


def test_plan_tiles(built_lib):
    lib = _lib.load()
    tile = lib.acb_frames_per_tile()
    lens = np.array([513, 8000, 16000, 480000, 4097], dtype=np.int64)
    ts = np.zeros(len(lens) + 1, dtype=np.int32)
    total = lib.acb_plan_tiles(lens.ctypes.data, len(lens), 1024, 256, 0, ts.ctypes.data)
    exp = [-(-(1 + int(n) // 256) // tile) for n in lens]
    assert total == sum(exp) and list(np.diff(ts)) == exp and ts[0] == 0
    total = lib.acb_plan_tiles(lens.ctypes.data, len(lens), 1024, 256, 1880, ts.ctypes.data)
    assert total == len(lens) * -(-1880 // tile)
    bad = np.array([8000, 512], dtype=np.int64)
    assert lib.acb_plan_tiles(bad.ctypes.data, 2, 1024, 256, 0, ts.ctypes.data) < 0
    assert b"clip 1" in lib.acb_last_error()
    assert lib.acb_plan_tiles(lens.ctypes.data, 0, 1024, 256, 0, ts.ctypes.data) == 0     # empty batch


def test_moments_finalize_matches_reference_algorithm(built_lib, golden, manifest):
    files = [golden[f"pipeline_noise_{n}_s{s}"] for n, s in ((16000, 1), (40000, 2), (100001, 3))]
    s, s2, frames = o.stats_per_bin(files)
    m = np.concatenate([s, s2])
    st = manifest["stats_three_files"]
    for fin in (stats.finalize_moments, stats.finalize_moments_c):
        r = fin(m, frames)
        assert r.count == st["total_count"] and r.frames == 616
        assert abs(r.mel_mean - st["mean"]) < 1e-6 and abs(r.mel_std - st["std"]) < 1e-6
        assert r.lines() == st["printed"]
        for i, b in enumerate((0, 40, 79)):
            assert abs(r.bin_mean[b] - st["per_bin_mean_0_40_79"][i]) < 1e-9
            assert abs(r.bin_std[b] - st["per_bin_std_0_40_79"][i]) < 1e-9
    # variance floor (compute_mel_stats.py:32)
    const = np.concatenate([np.full(80, 3.0 * 10), np.full(80, 9.0 * 10)])
    r = stats.finalize_moments(const, 10)
    assert np.allclose(r.bin_std, 1e-4) and abs(r.mel_std - 1e-4) < 1e-12


def test_sharding_partitions():
    for n, w in ((10, 4), (0, 3), (7, 8), (100000, 8)):
        seen = []
        for r in range(w):
            seen += list(sharding.contiguous_shard(n, r, w))
        assert seen == list(range(n))
    rng = np.random.default_rng(0)
    lens = rng.integers(16000, 480001, size=1000)
    shards = sharding.balanced_shards(lens, 8)
    allidx = np.sort(np.concatenate(shards))
    assert np.array_equal(allidx, np.arange(1000))
    loads = np.array([lens[s].sum() for s in shards])
    assert loads.max() / loads.mean() < 1.01
    batches = sharding.batches_by_budget(lens, shards[0], 4_000_000)
    assert np.array_equal(np.concatenate(batches), shards[0])
    assert all(lens[b].sum() <= 4_000_000 or len(b) == 1 for b in batches)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_product_package_never_touches_the_oracle_or_a_cpu_fallback():
    """The oracle is test infrastructure: nothing under audio-calm_b200/ may import it, and the arithmetic entry points must
    refuse CPU tensors instead of falling back."""
    import re
    pkg = os.path.join(ROOT, "audio-calm_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, re.M) or "ref_torch_port" in src or "logmel_oracle" in src:
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders
    import torch
    from audio_calm_b200.preprocess.core import MelExtractor
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MelExtractor()(torch.zeros(1, 4000))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        acb.LogMelFrontend("cpu")
