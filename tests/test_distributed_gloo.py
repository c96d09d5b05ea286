"""N>1 host logic on CPU: utterance sharding + the single all-reduce of the stats pass (gloo, world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import audio_calm_b200 as acb
    from audio_calm_b200 import sharding
    from oracle import logmel_oracle as o
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(ROOT, "tests", "golden", "cases.npz"))
    names = ["pipeline_noise_16000_s1", "pipeline_noise_40000_s2", "pipeline_noise_100001_s3", "pipeline_synth_24001_s12"]
    lens = [g[n].shape[1] for n in names]
    mine = sharding.balanced_shards(lens, world)[rank]
    acc = acb.MelStatsAccumulator(80, device="cpu")
    if len(mine):
        s, s2, frames = o.stats_per_bin([g[names[i]] for i in mine])       # per-rank moments (the GPU kernel's job on the box)
        acc.moments += torch.from_numpy(np.concatenate([s, s2]))
        acc.frames += frames
    calls = {"n": 0}
    real = dist.all_reduce

    def counting(*a, **k):
        calls["n"] += 1
        return real(*a, **k)
    dist.all_reduce = counting
    acc.all_reduce()
    dist.all_reduce = real
    st = acc.finalize()
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), mean=st.bin_mean, std=st.bin_std, g=np.array([st.mel_mean, st.mel_std]),
             frames=st.frames, calls=calls["n"], mine=np.array(mine))
    dist.destroy_process_group()


def test_stats_allreduce_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from oracle import logmel_oracle as o
    g = np.load(os.path.join(ROOT, "tests", "golden", "cases.npz"))
    names = ["pipeline_noise_16000_s1", "pipeline_noise_40000_s2", "pipeline_noise_100001_s3", "pipeline_synth_24001_s12"]
    s, s2, frames = o.stats_per_bin([g[n] for n in names])
    bm, bs = o.stats_per_bin_finalise(s, s2, frames)
    S, S2, N = o.stats_accumulate_scalar([g[n] for n in names])
    mean, std = o.stats_finalise(S, S2, N)
    r = [np.load(tmp_path / f"r{k}.npz") for k in range(world)]
    assert sorted(np.concatenate([x["mine"] for x in r]).tolist()) == [0, 1, 2, 3]      # disjoint cover of the utterances
    for x in r:
        assert int(x["calls"]) == 1                                                      # exactly one collective
        assert int(x["frames"]) == frames and frames * 80 == N
        assert np.max(np.abs(x["mean"] - bm)) < 1e-12 and np.max(np.abs(x["std"] - bs)) < 1e-12
        assert abs(x["g"][0] - mean) < 1e-6 and abs(x["g"][1] - std) < 1e-6
    assert np.array_equal(r[0]["mean"], r[1]["mean"])                                    # identical on every rank
