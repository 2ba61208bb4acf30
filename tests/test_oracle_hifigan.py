"""The HiFi-GAN oracle (oracle/hifigan_oracle.py) against outputs of the unmodified reference Generator
(tests/golden/hifigan_small.pt, written by oracle/make_goldens_hifigan.py), and -- in the authoring container, where
/root/reference exists -- against the reference run on its bundled trained weights.  CPU only."""
import os

import pytest
import torch

from lightningfastspeech2_b200 import synthetic
from oracle import hifigan_oracle as HO

REF = "/root/reference/litfass/third_party/hifigan"


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "hifigan_small.pt"), weights_only=False)


def test_oracle_reproduces_the_reference_generator(golden):
    sd = synthetic.hifigan_state_dict(golden["config"], seed=golden["seed"])
    assert sorted(sd) == golden["state_dict_keys"]          # the reference's key names (weight_norm removed)
    with torch.no_grad():
        wav = HO.generator(sd, golden["mel"], golden["config"])
        wav64 = HO.generator({k: v.double() for k, v in sd.items()}, golden["mel"].double(), golden["config"])
    assert wav.shape == golden["wav"].shape == (2, 1, 23 * 256)
    assert (wav - golden["wav"]).abs().max() < 1e-6
    assert (wav64 - golden["wav64"]).abs().max() < 1e-12
    for m, w in zip(golden["ragged_mels"], golden["ragged_wavs"]):
        with torch.no_grad():
            got = HO.generator(sd, m.T.unsqueeze(0), golden["config"])[0, 0]
        assert got.shape == (m.shape[0] * 256,) and (got - w).abs().max() < 1e-6


def test_synthesise_returns_int16_like_the_reference(golden):
    sd = synthetic.hifigan_state_dict(golden["config"], seed=golden["seed"])
    m = golden["ragged_mels"][1]
    out = HO.synthesise(sd, m, golden["config"])
    assert out.dtype.name == "int16" and out.shape == (1, m.shape[0] * 256)
    ref = (golden["ragged_wavs"][1].numpy() * 32768.0).astype("int16")
    assert abs(out[0].astype(int) - ref.astype(int)).max() <= 1


def test_fold_weight_norm_matches_torch():
    conv = torch.nn.utils.weight_norm(torch.nn.ConvTranspose1d(6, 4, 4, 2, padding=1))
    with torch.no_grad():
        conv.weight_g.mul_(1.7)
    sd = HO.fold_weight_norm({k: v.detach() for k, v in conv.state_dict().items()})
    want = torch._weight_norm(conv.weight_v, conv.weight_g, 0).detach()
    assert set(sd) == {"weight", "bias"} and (sd["weight"] - want).abs().max() < 1e-6


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "generator_universal.pth.tar")),
                    reason="the reference's bundled weights only exist in the authoring container")
def test_oracle_on_the_bundled_trained_weights(golden):
    ck = torch.load(os.path.join(REF, "generator_universal.pth.tar"), map_location="cpu", weights_only=False)["generator"]
    sd = HO.fold_weight_norm(ck)
    with torch.no_grad():
        wav = HO.generator(sd, golden["mel"][:1] * 2 - 4, golden["config"])[0, 0]
    assert (wav - golden["universal_wav_first_mel"]).abs().max() < 2e-6
