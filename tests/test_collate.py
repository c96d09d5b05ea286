"""Training-feed collation: the device kernels against the reference's host code restated with torch ops
(train/train_vae.py:83-116 crop / zero-pad + stack; train/train_calm.py:184-215 SpecAugment span + pad_sequence + transpose)."""
import numpy as np
import pytest
import torch

import audio_calm_b200 as acb
from audio_calm_b200 import collate


def ref_crop(mel, crop, start):
    """MelDataset.__getitem__ for one item with a given start (train_vae.py:86-102)."""
    if mel.shape[1] > crop:
        return mel[:, start:start + crop]
    return torch.nn.functional.pad(mel, (0, crop - mel.shape[1]))


def test_crop_starts_follow_the_reference_rules():
    frames = torch.tensor([10, 256, 257, 300, 1000])
    ev = collate.crop_starts(frames, 256, is_eval=True)
    assert ev.tolist() == [0, 0, 0, 22, 372]                                      # (T - crop) // 2, 0 when T <= crop
    g = torch.Generator().manual_seed(0)
    for _ in range(50):
        st = collate.crop_starts(frames, 256, is_eval=False, generator=g)
        assert st[0] == 0 and st[1] == 0 and st[2] == 0                          # randint(0, 1) is always 0
        assert 0 <= st[3] < 44 and 0 <= st[4] < 744                              # randint upper bound is exclusive


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_crop_collate_matches_reference(dtype):
    g = torch.Generator().manual_seed(1)
    frames = torch.tensor([40, 256, 300, 777, 1024, 5])
    cap = 1024
    feat = torch.randn(6, 80, cap, generator=g).to(dtype)
    for i, n in enumerate(frames.tolist()):
        feat[i, :, n:] = 99.0                                                     # garbage beyond the valid frames must not leak
    for is_eval in (True, False):
        out = acb.crop_collate(feat.cuda(), frames, crop_size=256, is_eval=is_eval, generator=g)
        assert out["mel"] is out["labels"] and tuple(out["mel"].shape) == (6, 80, 256) and out["mel"].dtype == dtype
        st = out["start"].cpu().tolist()
        for i, n in enumerate(frames.tolist()):
            ref = ref_crop(feat[i, :, :n], 256, st[i])
            assert torch.equal(out["mel"][i].cpu(), ref), (i, is_eval)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pad_collate_matches_calm_collator(dtype):
    g = torch.Generator().manual_seed(2)
    lens = [37, 1, 384, 100, 65]
    items = [torch.randn(n, 128, generator=g).to(dtype) for n in lens]
    out = acb.pad_collate(items, pad_value=0.0)
    ref = torch.nn.utils.rnn.pad_sequence(items, batch_first=True, padding_value=0.0).transpose(1, 2)   # train_calm.py:212-214
    assert out["audio_lens"].tolist() == lens and out["audio_lens"].dtype == torch.int64
    assert torch.equal(out["audio_features"].cpu(), ref.contiguous())
    # SpecAugment span (train_calm.py:184-191) on some clips, a different pad value, a wider batch
    m0 = torch.tensor([3, 0, 100, 0, 60]); m1 = torch.tensor([7, 0, 10, 0, 5])
    out2 = acb.pad_collate(items, pad_value=-1.0, mask=(m0, m1), out_frames=400)
    masked = []
    for x, a, n in zip(items, m0.tolist(), m1.tolist()):
        y = x.clone()
        y[a:a + n] = 0.0
        masked.append(y)
    ref2 = torch.full((5, 128, 400), -1.0, dtype=dtype)
    for i, y in enumerate(masked):
        ref2[i, :, :y.shape[0]] = y.transpose(0, 1)
    assert torch.equal(out2["audio_features"].cpu(), ref2)


@pytest.mark.gpu
def test_feature_batch_to_vae_crop():
    """End of the chain: ragged waveforms -> padded log-mel batch -> 256-frame training crops, all on the device."""
    from oracle import logmel_oracle as o
    fe = acb.LogMelFrontend("cuda")
    clips = [o.synth_clip(n, 800 + i) for i, n in enumerate([30000, 70001, 131072])]
    batch = acb.pack_clips([torch.from_numpy(c) for c in clips], fe.device)
    feats, frames = fe.forward_ragged(batch, pad_multiple=4, peak=fe.peak_abs_ragged(batch))
    out = acb.crop_collate(feats, frames, crop_size=256, is_eval=True)
    for i, c in enumerate(clips):
        ref = torch.from_numpy(o.dataset_mel(c[None], fe.window.numpy(), fe.fb.numpy()).astype(np.float32))
        want = ref_crop(ref, 256, (ref.shape[1] - 256) // 2 if ref.shape[1] > 256 else 0)
        assert float((out["mel"][i].cpu() - want).abs().max()) < 1e-4
