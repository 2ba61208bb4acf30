"""The FastDiff variance adaptor oracle (oracle/fastdiff_oracle.py) against recorded runs of the unmodified reference
module (tests/golden/fastdiff_adaptor.pt, written by oracle/make_goldens_fastdiff.py).  CPU only."""
import os

import pytest
import torch

from lightningfastspeech2_b200 import synthetic
from oracle import fastdiff_oracle as FO


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "fastdiff_adaptor.pt"), weights_only=False)


def _sd(g):
    return synthetic.fill_state_dict({k: None for k in g["state_dict_keys"]} and _shapes(g), seed=g["seed"])


def _shapes(g):
    """state_dict of this repo's mirror = the reference's keys (checked below) with the reference's shapes"""
    from lightningfastspeech2_b200.fastspeech2.fastdiff_variances import FastDiffVarianceAdaptor

    c = g["cfg"]
    ada = FastDiffVarianceAdaptor(c["stats"], c["variances"], c["variance_nlayers"], c["variance_kernel_size"],
                                  c["variance_dropout"], c["variance_filter_size"], c["variance_nbins"],
                                  c["variance_depthwise_conv"], c["duration_nlayers"], c["duration_kernel_size"],
                                  c["duration_dropout"], c["duration_filter_size"], c["duration_depthwise_conv"],
                                  c["encoder_hidden"], c["max_length"])
    sd = ada.state_dict()
    assert sorted(sd) == g["state_dict_keys"]
    return sd


def test_inference_reproduces_the_reference(golden):
    sd = _sd(golden)
    ref = golden["inference"]["out"]
    with torch.no_grad():
        r = FO.adaptor(sd, golden["cfg"], golden["x"], golden["src_mask"], {}, True, golden["inference"]["noise"])
    assert torch.equal(r["duration_rounded"], ref["duration_rounded"]) and torch.equal(r["tgt_mask"], ref["tgt_mask"])
    assert r["x"].shape == ref["x"].shape and r["x"].shape[1] % 64 == 0
    for k in ("duration_prediction", "variances_pitch", "variances_energy", "x", "out"):
        assert (r[k] - ref[k]).abs().max() < 2e-5, k


def test_teacher_forced_reproduces_the_reference(golden):
    sd = _sd(golden)
    tf = golden["teacher_forced"]
    steps = dict(zip(["duration"] + golden["cfg"]["variances"], tf["steps"]))
    with torch.no_grad():
        r = FO.adaptor(sd, golden["cfg"], golden["x"], golden["src_mask"], tf["targets"], False, tf["noise"], steps=steps,
                       jitter=tf["jitter"])
    ref = tf["out"]
    for k in ("duration_prediction", "variances_pitch", "variances_energy", "x", "out", "duration_z", "variances_pitch_z"):
        assert (r[k] - ref[k]).abs().max() < 2e-5, k
    assert torch.equal(r["tgt_mask"], ref["tgt_mask"])


def test_length_regulator_pad_to_multiple(golden):
    x = golden["x"]
    dur = golden["inference"]["out"]["duration_rounded"]
    out, mask = FO.length_regulator_padded(x, dur, 2756.25, 64)
    assert out.shape[1] % 64 == 0 and torch.equal(mask, golden["inference"]["out"]["tgt_mask"])
