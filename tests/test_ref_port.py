"""The torch CPU port used for the CPU baseline must reproduce the reference's golden outputs bit-for-bit
(same torch/torchaudio build) or within fp32 round-off (different build)."""
import numpy as np
import torch

from oracle import logmel_oracle as o
from oracle.ref_torch_port import RefMelExtractor, normalise, process_audio_chunk


def test_port_matches_golden(golden, manifest):
    torch.set_num_threads(1)
    ext = RefMelExtractor().eval()
    same_build = torch.__version__ == manifest["torch"]
    with torch.inference_mode():
        for name, x in (("raw_noise_16000_s1", o.hash_noise(16000, 1)), ("raw_synth_24001_s12", o.synth_clip(24001, 12)),
                        ("raw_synth_513_s13", o.synth_clip(513, 13))):
            y = ext(torch.from_numpy(x)[None])[0].numpy()
            d = np.max(np.abs(y - golden[name]))
            assert d == 0.0 if same_build else d < 5e-6
        st = np.stack([o.hash_noise(5000, 31), o.synth_clip(5000, 32)])
        assert np.array_equal(process_audio_chunk(torch.from_numpy(st)).numpy(), golden["chunk_stereo_5000"])
        m = torch.from_numpy(golden["pipeline_noise_16000_s1"])
        assert np.max(np.abs(normalise(m).numpy() - golden["norm_global_noise_16000_s1"])) < 1e-6
