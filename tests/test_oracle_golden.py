"""The oracle is only trusted after it reproduces every golden vector minted from the reference."""
import numpy as np
import pytest

from oracle import logmel_oracle as o

NOISE = [(16000, 1), (40000, 2), (100001, 3)]
SYNTH = [(8000, 11), (24001, 12), (513, 13), (1024, 14), (777, 15), (1279, 16), (1280, 17)]
TOL = 2e-6   # fp64 oracle vs the fp32 reference on noise-like input (measured <= 8.6e-7)


def _tone():
    t = np.arange(16000, dtype=np.float64) / 16000.0
    return (0.5 * np.sin(2 * np.pi * 440 * t) + 0.25 * np.sin(2 * np.pi * 3000 * t + 1.0)).astype(np.float32)


@pytest.mark.parametrize("n,seed", NOISE)
def test_pipeline_noise(golden, golden_tables, n, seed):
    ref = golden[f"pipeline_noise_{n}_s{seed}"]
    got = o.dataset_mel(o.hash_noise(n, seed)[None], golden_tables["window"], golden_tables["fb"])
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref)) < TOL
    got32 = o.dataset_mel(o.hash_noise(n, seed)[None], golden_tables["window"], golden_tables["fb"], dtype=np.float32)
    assert np.max(np.abs(got32 - ref)) < 5e-6


@pytest.mark.parametrize("n,seed", SYNTH)
def test_raw_synth(golden, golden_tables, n, seed):
    ref = golden[f"raw_synth_{n}_s{seed}"]
    got = o.logmel(o.synth_clip(n, seed), golden_tables["window"], golden_tables["fb"])
    assert got.shape == ref.shape == (80, 1 + n // 256)
    assert np.max(np.abs(got - ref)) < TOL


def test_raw_noise_and_pipeline_synth(golden, golden_tables):
    w, fb = golden_tables["window"], golden_tables["fb"]
    assert np.max(np.abs(o.logmel(o.hash_noise(16000, 1), w, fb) - golden["raw_noise_16000_s1"])) < TOL
    assert np.max(np.abs(o.dataset_mel(o.synth_clip(24001, 12)[None], w, fb) - golden["pipeline_synth_24001_s12"])) < TOL


def test_fp64_tables_shift(golden):
    """With fp64-derived tables the oracle moves by up to a few 1e-5 -- still far inside 1e-4."""
    got = o.dataset_mel(o.hash_noise(16000, 1)[None])
    d = np.max(np.abs(got - golden["pipeline_noise_16000_s1"]))
    assert 1e-6 < d < 5e-5


def test_tone_clamp_and_zeros(golden, golden_tables):
    w, fb = golden_tables["window"], golden_tables["fb"]
    ref = golden["raw_tone_16000"]
    got = o.logmel(_tone(), w, fb)
    floor = np.float32(o.LOG_FLOOR)
    assert ref[5, 10] == floor and ref[79, 10] == floor
    # tonal, high dynamic range: the reference's own fp32 error dominates; only a loose bound holds
    assert np.max(np.abs(got - ref)) < 3e-4
    z = o.logmel(np.zeros(4000, np.float32), w, fb)
    assert np.all(golden["raw_zeros_4000"] == floor)
    assert np.max(np.abs(z - golden["raw_zeros_4000"])) < 1e-6


def test_batch(golden, golden_tables):
    w, fb = golden_tables["window"], golden_tables["fb"]
    clips = [o.synth_clip(12000, 21), o.synth_clip(12000, 22), o.hash_noise(12000, 23)]
    ref = golden["raw_batch3_12000"]
    for i, c in enumerate(clips):
        assert np.max(np.abs(o.logmel(c, w, fb) - ref[i])) < TOL


def test_process_audio_chunk_bitwise(golden):
    st = np.stack([o.hash_noise(5000, 31), o.synth_clip(5000, 32)])
    assert np.array_equal(o.process_audio_chunk(st), golden["chunk_stereo_5000"])
    assert np.array_equal(o.process_audio_chunk(o.synth_clip(5000, 33)[None]), golden["chunk_mono_5000"])
    assert np.array_equal(o.process_audio_chunk(np.zeros((1, 100), np.float32)), golden["chunk_zeros_100"])
    assert abs(float(np.abs(golden["chunk_mono_5000"]).max()) - 0.95) < 1e-6


def test_pad_reflect_and_normalisations(golden):
    m1 = golden["pipeline_noise_16000_s1"]
    assert m1.shape == (80, 64) and np.array_equal(m1[:, 63], m1[:, 61])          # T=63 -> one reflected column
    m3 = golden["pipeline_noise_100001_s3"]
    assert m3.shape == (80, 392) and np.array_equal(m3[:, 391], m3[:, 389])       # T=391 -> one reflected column
    m2 = golden["pipeline_noise_40000_s2"]
    assert m2.shape == (80, 160)                                                   # T=157 -> 3 columns
    for j in range(3):
        assert np.array_equal(m2[:, 157 + j], m2[:, 155 - j])
    x = np.arange(10, dtype=np.float64)[None, :7]
    assert np.array_equal(o.pad_time_reflect(x, 4)[0], [0, 1, 2, 3, 4, 5, 6, 5])
    assert np.max(np.abs(o.normalise_global(m1.astype(np.float64)) - golden["norm_global_noise_16000_s1"])) < 1e-6
    assert np.max(np.abs(o.normalise_per_utterance(m1.astype(np.float64)) - golden["norm_utt_noise_16000_s1"])) < 2e-6


def test_stats_pass(golden, manifest):
    files = [golden[f"pipeline_noise_{n}_s{s}"] for n, s in NOISE]
    st = manifest["stats_three_files"]
    S, S2, N = o.stats_accumulate_scalar(files)
    assert N == st["total_count"] == 49280
    assert abs(S - st["total_sum"]) / abs(st["total_sum"]) < 1e-6
    assert abs(S2 - st["total_sq_sum"]) / st["total_sq_sum"] < 1e-6
    mean, std = o.stats_finalise(S, S2, N)
    assert abs(mean - st["mean"]) < 1e-6 and abs(std - st["std"]) < 1e-6
    assert list(o.format_stats_lines(mean, std)) == st["printed"]
    s, s2, frames = o.stats_per_bin(files)
    assert frames * 80 == N
    bm, bs = o.stats_per_bin_finalise(s, s2, frames)
    for i, b in enumerate((0, 40, 79)):
        assert abs(bm[b] - st["per_bin_mean_0_40_79"][i]) < 1e-9
        assert abs(bs[b] - st["per_bin_std_0_40_79"][i]) < 1e-9
    assert abs(bm.mean() - mean) < 1e-6                                             # mean of per-bin means = scalar mean


def test_frame_table_and_errors(manifest):
    for L, T in manifest["frame_table"].items():
        assert o.frames_for_length(int(L)) == T
    for L, err in manifest["short_input_error"].items():
        assert err == "RuntimeError"
        with pytest.raises(RuntimeError):
            o.frames_for_length(int(L))
    with pytest.raises(RuntimeError):
        o.logmel(np.zeros(512, np.float32))


def test_hash_noise_known_answer():
    x = o.hash_noise(16000, 1)
    assert np.allclose(x[:4], [0.26630175, -0.37396902, 0.20093119, 0.13287622], atol=1e-8)
    assert abs(float(np.abs(x).max()) - 0.49997401) < 1e-7


def test_oracle_matches_the_benchmark_distribution_goldens():
    """configs[1] at its own size: 30 s clips of the benchmark's signal through the reference (oracle/gen_golden_bench.py)."""
    import os
    import audio_calm_b200 as acb
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bench_cases.npz"))
    w, fb = acb.tables.calm_tables()
    x = o.bench_clip(480000, 202)
    y = o.normalise_global(o.logmel(x, w.numpy(), fb.numpy()))
    assert y.shape == (80, 1876)
    assert float(np.max(np.abs(y[:, ::7] - g["norm_bench_30s_s202_sub7"]))) < 1e-5
    mn, mx, mean = g["norm_bench_30s_s202_stats"]
    assert abs(y.min() - mn) < 1e-5 and abs(y.max() - mx) < 1e-5 and abs(float(y.mean()) - mean) < 1e-6
