"""The dataset driver mirror (preprocess/process_dataset.py) and the statistics CLIs around the hot path.

CPU tests cover the host-side bookkeeping (paths, resume, transcripts, sharding, CLI flags); the GPU tests run the driver end to end on
synthetic .wav files and compare every saved payload with the oracle's restatement of process_dataset.py:140-156."""
import os
import types

import numpy as np
import pytest
import torch

import audio_calm_b200 as acb
from audio_calm_b200.preprocess import compute_latent_stats, compute_mel_stats, process_dataset as pd
from oracle import logmel_oracle as o


def _args(**kw):
    base = dict(dataset_name="librispeech", in_dir="", out_dir="", vae_ckpt=None, mel_only=True, cv_tsv=None, num_gpus=1,
                workers_per_gpu=2, force=False)
    base.update(kw)
    return types.SimpleNamespace(**base)


def _write_wav(path, x):
    from scipy.io import wavfile
    os.makedirs(os.path.dirname(path), exist_ok=True)
    wavfile.write(path, 16000, x.astype(np.float32))


def test_cli_flags_match_the_reference():
    a = pd.build_parser().parse_args(["--dataset_name", "libritts", "--in_dir", "i", "--out_dir", "o", "--mel_only", "--force",
                                      "--num_gpus", "2", "--workers_per_gpu", "3", "--cv_tsv", "x.tsv", "--vae_ckpt", "c"])
    assert (a.dataset_name, a.in_dir, a.out_dir, a.mel_only, a.force, a.num_gpus, a.workers_per_gpu, a.cv_tsv, a.vae_ckpt) == \
           ("libritts", "i", "o", True, True, 2, 3, "x.tsv", "c")
    assert a.procs_per_gpu == 1                                                           # extension flag: one process per GPU unless asked


def test_scan_and_output_paths(tmp_path):
    root = tmp_path / "in"
    for rel in ("a/b/u1.wav", "a/b/u2.flac", "a/c/u3.mp3", "a/c/notes.txt"):
        p = root / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_bytes(b"")
    files = pd.scan_files(str(root))
    assert sorted(os.path.basename(f) for f in files) == ["u1.wav", "u2.flac", "u3.mp3"]
    args = _args(in_dir=str(root), out_dir=str(tmp_path / "out"))
    save_dir, file_id, save_path = pd.output_path(str(root / "a/b/u1.wav"), args)
    assert save_dir == str(tmp_path / "out" / "a" / "b") and file_id == "u1" and save_path.endswith("a/b/u1.pt")
    args.dataset_name = "commonvoice"                       # flat output (process_dataset.py:113-116)
    assert pd.output_path(str(root / "a/b/u1.wav"), args)[0] == str(tmp_path / "out")


def test_transcripts(tmp_path):
    d = tmp_path / "spk" / "chap"
    d.mkdir(parents=True)
    (d / "spk-chap.trans.txt").write_text("spk-chap-0001 HELLO WORLD\nspk-chap-0002 SECOND LINE\n")
    (d / "utt.normalized.txt").write_text(" normalised text \n")
    assert pd.transcript_for(str(d / "spk-chap-0002.flac"), _args(dataset_name="librispeech"), {}) == "SECOND LINE"
    assert pd.transcript_for(str(d / "utt.wav"), _args(dataset_name="libritts"), {}) == "normalised text"
    assert pd.transcript_for(str(d / "x.mp3"), _args(dataset_name="commonvoice"), {"x.mp3": "bonjour"}) == "bonjour"
    tsv = tmp_path / "cv.tsv"
    tsv.write_text("path\tsentence\nx.mp3\tbonjour\n")
    assert pd.get_common_voice_map(str(tsv)) == {"x.mp3": "bonjour"}


def test_load_audio_pcm_wav(tmp_path):
    from scipy.io import wavfile
    x = (o.hash_noise(4000, 3) * 30000).astype(np.int16)
    wavfile.write(str(tmp_path / "a.wav"), 16000, x)
    w = pd.load_audio(str(tmp_path / "a.wav"))
    assert tuple(w.shape) == (1, 4000)
    if w.dtype == torch.int16:                                # 16 kHz 16-bit files travel as PCM and are widened on the device
        assert np.array_equal(w.numpy()[0], x)
    else:
        assert w.dtype == torch.float32 and np.allclose(w.numpy()[0], x.astype(np.float32) / 32768.0, atol=1e-7)
    wavfile.write(str(tmp_path / "b.wav"), 16000, (x.astype(np.float32) / 32768.0))
    wf = pd.load_audio(str(tmp_path / "b.wav"))
    assert wf.dtype == torch.float32 and np.allclose(wf.numpy()[0], x.astype(np.float32) / 32768.0, atol=1e-7)


def test_iter_files_generators(tmp_path):
    for rel in ("x/a.pt", "x/y/b.pt", "x/c.txt"):
        p = tmp_path / rel
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_bytes(b"")
    assert sorted(os.path.basename(p) for p in compute_mel_stats.iter_mel_files(str(tmp_path))) == ["a.pt", "b.pt"]


def test_failed_saves_are_reported_and_keep_their_transcript_out(tmp_path):
    """A torch.save / makedirs failure must surface in runner.errors, and the clip's transcript line must not be written
    (process_dataset.py:152-189 appends the entry only after the save)."""
    from concurrent.futures import ThreadPoolExecutor
    runner = pd.ShardRunner.__new__(pd.ShardRunner)            # host bookkeeping only: no device objects needed
    runner.args, runner.errors, runner.trans_buffer, runner.done = _args(out_dir=str(tmp_path)), [], {}, 0
    runner.report = lambda n: None
    blocked = tmp_path / "blocked"
    blocked.write_text("a file where a directory is needed")   # makedirs(save_dir) fails under it
    good_dir = str(tmp_path / "spk" / "chap")
    ok = pd.Clip("in/spk/chap/u1.flac", good_dir, "u1", os.path.join(good_dir, "u1.pt"), "HELLO", torch.zeros(1, 1))
    bad = pd.Clip("in/x/u2.flac", str(blocked / "x"), "u2", str(blocked / "x" / "u2.pt"), "LOST", torch.zeros(1, 1))
    with ThreadPoolExecutor(1) as ex:
        runner._inflight = [(ex.submit(pd.ShardRunner._save, c.save_dir, c.save_path, {"mel": torch.zeros(80, 4)}), c) for c in (ok, bad)]
        runner._collect(wait=True)
    assert os.path.exists(ok.save_path) and not runner._inflight
    assert len(runner.errors) == 1 and runner.errors[0][0] == bad.wav_path and "save failed" in runner.errors[0][1]
    assert runner.trans_buffer == {os.path.join(good_dir, "chap.trans.txt"): ["u1 HELLO"]}
    assert runner.done == 2                                      # both files count as processed for the progress bar


def test_mel_stats_file_round_trip(tmp_path):
    """MelStats.save -> MelStats.load keeps the per-bin vectors and the scalar pair; the reference's bare {"mean","std"}
    layout (compute_latent_stats.py:44-47) loads too."""
    rng = np.random.default_rng(3)
    s, frames = rng.normal(-500.0, 20.0, 80), 100
    s2 = s * s / frames + rng.uniform(50.0, 90.0, 80)
    st = acb.finalize_moments(np.concatenate([s, s2]), frames)
    path = str(tmp_path / "mel_stats.pt")
    st.save(path)
    back = acb.MelStats.load(path)
    assert np.allclose(back.bin_mean, st.bin_mean, rtol=0, atol=1e-6 * np.abs(st.bin_mean).max())     # stored as fp32
    assert np.allclose(back.bin_std, st.bin_std, rtol=1e-6)
    assert back.mel_mean == st.mel_mean and back.mel_std == st.mel_std and back.frames == frames and back.count == 80 * frames
    mean_t, std_t = back.affine()
    assert mean_t.dtype == torch.float32 and tuple(mean_t.shape) == (80,) and tuple(std_t.shape) == (80,)
    assert back.affine(per_bin=False) == (st.mel_mean, st.mel_std)
    torch.save({"mean": torch.from_numpy(st.bin_mean.astype(np.float32)), "std": torch.from_numpy(st.bin_std.astype(np.float32))},
               str(tmp_path / "bare.pt"))
    bare = acb.MelStats.load(str(tmp_path / "bare.pt"))
    assert abs(bare.mel_mean - st.mel_mean) < 1e-5 and abs(bare.mel_std - st.mel_std) < 1e-4
    with pytest.raises(ValueError):
        torch.save({"latent": torch.zeros(3)}, str(tmp_path / "no.pt"))
        acb.MelStats.load(str(tmp_path / "no.pt"))


# ------------------------------------------------------------------------------------------------ GPU: end to end
@pytest.mark.gpu
def test_driver_mel_only_matches_reference_pipeline(tmp_path):
    fe_tables = acb.tables.calm_tables()
    window, fb = fe_tables[0].numpy(), fe_tables[1].numpy()
    root, out = tmp_path / "in", tmp_path / "out"
    clips = {"s1/c1/u1.wav": o.synth_clip(16000, 21), "s1/c1/u2.wav": o.synth_clip(40001, 22), "s1/c2/u3.wav": o.hash_noise(24000, 23),
             "s2/c3/u4.wav": np.stack([o.synth_clip(20000, 24), o.hash_noise(20000, 25)], axis=1)}   # u4 is stereo
    for rel, x in clips.items():
        _write_wav(str(root / rel), x)
    _write_wav(str(root / "s2/c3/short.wav"), o.hash_noise(300, 26))                                  # too short: reported, not fatal
    from scipy.io import wavfile
    pcm = (o.synth_clip(30000, 27) * 32767).astype(np.int16)                                          # a 16-bit PCM file: shipped as int16
    wavfile.write(str(root / "s2/c3/u5.wav"), 16000, pcm)
    clips["s2/c3/u5.wav"] = pcm.astype(np.float32) / 32768.0
    args = _args(in_dir=str(root), out_dir=str(out))
    runner = pd.ShardRunner(args, 0, batch_samples=70000, decode_threads=2)                           # forces several ragged launches
    runner.run(pd.scan_files(str(root)))
    assert len(runner.errors) == 1 and "short.wav" in runner.errors[0][0]
    for rel, x in clips.items():
        payload = torch.load(str(out / rel.replace(".wav", ".pt")))
        mel = payload["mel"]
        mono = x if x.ndim == 1 else x.T                                                              # [C, L] for the oracle
        ref = o.dataset_mel(np.atleast_2d(mono), window, fb)
        assert mel.dtype == torch.float32 and tuple(mel.shape) == ref.shape and mel.shape[1] % 4 == 0
        assert float(np.max(np.abs(mel.numpy() - ref))) < 1e-4
    # resume: nothing is rewritten unless --force (process_dataset.py:125-130)
    stamp = {rel: os.path.getmtime(str(out / rel.replace(".wav", ".pt"))) for rel in clips}
    pd.ShardRunner(args, 0).run(pd.scan_files(str(root)))
    assert stamp == {rel: os.path.getmtime(str(out / rel.replace(".wav", ".pt"))) for rel in clips}


@pytest.mark.gpu
def test_cli_with_two_processes_on_one_gpu(tmp_path, capsys):
    """The command line end to end (spawned workers, progress queue, exit code) with --procs_per_gpu 2: contiguous shards over
    num_gpus * procs_per_gpu processes, rank % num_gpus picks the device (process_dataset.py:256-275)."""
    fe_tables = acb.tables.calm_tables()
    window, fb = fe_tables[0].numpy(), fe_tables[1].numpy()
    root, out = tmp_path / "in", tmp_path / "out"
    clips = {f"s{i % 2}/c{i % 3}/u{i}.wav": o.synth_clip(12000 + 997 * i, 40 + i) for i in range(7)}
    for rel, x in clips.items():
        _write_wav(str(root / rel), x)
    rc = pd.main(["--dataset_name", "librispeech", "--in_dir", str(root), "--out_dir", str(out), "--mel_only", "--num_gpus", "1",
                  "--procs_per_gpu", "2", "--workers_per_gpu", "2"])
    assert rc == 0 and "Done: 7/7 files" in capsys.readouterr().out
    for rel, x in clips.items():
        mel = torch.load(str(out / rel.replace(".wav", ".pt")))["mel"]
        assert float(np.max(np.abs(mel.numpy() - o.dataset_mel(x[None], window, fb)))) < 1e-4


@pytest.mark.gpu
def test_stored_stats_round_trip_normalises_like_the_oracle(tmp_path, capsys):
    """stats CLI --save -> MelStats.load -> LogMelFrontend.forward(affine=stats): "per-bin normalisation with the stored mel
    stats" end to end, against the oracle's (mel - mean[b]) / std[b] with the reference's statistics algorithm per bin."""
    fe = acb.LogMelFrontend("cuda")
    window, fb = fe.window.numpy(), fe.fb.numpy()
    waves = [o.hash_noise(16000, 1), o.synth_clip(40000, 2), o.hash_noise(100001, 3)]
    mels = [o.dataset_mel(w[None], window, fb).astype(np.float32) for w in waves]
    for i, m in enumerate(mels):
        torch.save({"mel": torch.from_numpy(m)}, str(tmp_path / f"m{i}.pt"))
    path = str(tmp_path / "mel_stats.pt")
    compute_mel_stats.main(["--root", str(tmp_path), "--save", path])
    capsys.readouterr()
    st = acb.MelStats.load(path)
    s, s2, frames = o.stats_per_bin(mels)
    bm, bs = o.stats_per_bin_finalise(s, s2, frames)
    assert st.frames == frames and np.max(np.abs(st.bin_mean - bm)) < 1e-5 and np.max(np.abs(st.bin_std - bs)) < 1e-5
    x = o.synth_clip(30000, 9)
    y = fe.forward(torch.from_numpy(x)[None].cuda(), affine=st.affine())[0].cpu().numpy()
    ref = (o.logmel(x, window, fb) - bm[:, None]) / bs[:, None]
    assert float(np.max(np.abs(y - ref))) < 1e-4
    y1 = fe.forward(torch.from_numpy(x)[None].cuda(), affine=st.affine(per_bin=False))[0].cpu().numpy()     # the reference's scalar pair
    assert float(np.max(np.abs(y1 - o.normalise_global(o.logmel(x, window, fb), st.mel_mean, st.mel_std)))) < 1e-4


@pytest.mark.gpu
def test_stats_cli_over_saved_features(tmp_path, capsys, manifest):
    fe_tables = acb.tables.calm_tables()
    window, fb = fe_tables[0].numpy(), fe_tables[1].numpy()
    mels = [o.dataset_mel(o.hash_noise(n, s)[None], window, fb).astype(np.float32) for n, s in ((16000, 1), (40000, 2), (100001, 3))]
    for i, m in enumerate(mels):
        torch.save({"mel": torch.from_numpy(m)}, str(tmp_path / f"m{i}.pt"))
    stats = compute_mel_stats.main(["--root", str(tmp_path)])
    lines = capsys.readouterr().out.strip().splitlines()
    want = manifest["stats_three_files"]                       # what the reference's compute_mel_stats.py printed for these files
    assert lines[-2:] == want["printed"]
    assert stats.count == want["total_count"]
    assert abs(stats.mel_mean - want["mean"]) < 1e-6 and abs(stats.mel_std - want["std"]) < 1e-6


def _ref_latent_stats(files, reduce_dim):
    """The reference loop (preprocess/compute_latent_stats.py:15-40) restated with the same torch calls, in fp64 to serve as truth."""
    sum1 = sum2 = None
    count = 0
    for f in files:
        payload = torch.load(f, map_location="cpu", weights_only=True)
        lat = payload.get("latent", payload) if isinstance(payload, dict) else payload
        if lat.dim() == 2 and lat.shape[0] in (64, 80, 128, 192):
            lat = lat.transpose(0, 1)
        lat = lat.double()
        lat = lat.contiguous().reshape(-1) if reduce_dim else lat.contiguous()
        s1, s2 = lat.sum(dim=0), (lat * lat).sum(dim=0)
        sum1, sum2 = (s1, s2) if sum1 is None else (sum1 + s1, sum2 + s2)
        count += lat.shape[0]
    mean = sum1 / count
    std = torch.sqrt((sum2 / count - mean * mean).clamp(min=1e-12))
    return mean, std


@pytest.mark.gpu
def test_latent_stats_scalar_and_per_dim(tmp_path, capsys):
    g = torch.Generator().manual_seed(5)
    files = []
    for i, (T, as_dt) in enumerate([(96, True), (41, False), (300, True), (7, False)]):      # (D, T) and (T, D) payloads, bare tensor too
        lat = torch.randn(T, 128, generator=g) * 1.2 + 0.04
        payload = {"latent": lat.transpose(0, 1).contiguous() if as_dt else lat, "vae_path": "x"} if i != 3 else lat
        p = tmp_path / "lat" / f"sub{i % 2}" / f"l{i}.pt"
        p.parent.mkdir(parents=True, exist_ok=True)
        torch.save(payload, str(p))
        files.append(str(p))
    files = sorted(files)
    mean, std = compute_latent_stats.latent_stats(files, "cuda", reduce_dim=True)
    rm, rs = _ref_latent_stats(files, True)
    assert abs(mean - float(rm)) < 1e-6 and abs(std - float(rs)) < 1e-6
    mean_d, std_d = compute_latent_stats.latent_stats(files, "cuda", reduce_dim=False)
    rm, rs = _ref_latent_stats(files, False)
    assert tuple(mean_d.shape) == (128,) and mean_d.dtype == torch.float32
    assert float((mean_d.double() - rm).abs().max()) < 1e-6 and float((std_d.double() - rs).abs().max()) < 1e-6
    out = tmp_path / "latent_stats.pt"
    compute_latent_stats.main(["--latent_dir", str(tmp_path / "lat"), "--per_dim", "--out", str(out)])
    saved = torch.load(str(out))
    assert set(saved) == {"mean", "std"} and tuple(saved["mean"].shape) == (128,)           # the reference's only on-disk stats layout
    compute_latent_stats.main(["--latent_dir", str(tmp_path / "lat")])
    lines = capsys.readouterr().out.strip().splitlines()
    assert lines[-2] == f"latent_mean: {mean:.6f}" and lines[-1] == f"latent_std : {std:.6f}"
