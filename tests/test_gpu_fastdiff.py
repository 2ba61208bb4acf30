"""FastDiff variance adaptor on the GPU against recorded runs of the reference module
(tests/golden/fastdiff_adaptor.pt) with the reference's random draws injected, and the pad_to_multiple_of LengthRegulator.
Tolerances: continuous outputs 1e-3 abs in "fp32" mode (the reverse diffusion amplifies by up to 1.8x per step);
durations / masks exact (the golden's durations are forced when a rounding decision flips)."""
import os

import pytest
import torch

from lightningfastspeech2_b200 import configs, ops, synthetic
from lightningfastspeech2_b200.fastspeech2.fastdiff_variances import FastDiffVarianceAdaptor
from lightningfastspeech2_b200.fastspeech2.fastspeech2 import FastSpeech2
from lightningfastspeech2_b200.fastspeech2.model import LengthRegulator

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "fastdiff_adaptor.pt"), weights_only=False)


def build(g, mode="fp32"):
    c = g["cfg"]
    ada = FastDiffVarianceAdaptor(c["stats"], c["variances"], c["variance_nlayers"], c["variance_kernel_size"],
                                  c["variance_dropout"], c["variance_filter_size"], c["variance_nbins"],
                                  c["variance_depthwise_conv"], c["duration_nlayers"], c["duration_kernel_size"],
                                  c["duration_dropout"], c["duration_filter_size"], c["duration_depthwise_conv"],
                                  c["encoder_hidden"], c["max_length"])
    ada.load_state_dict(synthetic.fill_state_dict(ada.state_dict(), seed=g["seed"]))
    ada = ada.eval().to(DEV)
    for m in ada.modules():
        if hasattr(type(m), "compute_mode"):
            m.compute_mode = mode
    return ada


def _buckets(g, var, pred):
    st = g["cfg"]["stats"][var]
    bins = torch.linspace(st["min"], st["max"], g["cfg"]["variance_nbins"] - 1)
    return torch.bucketize(pred.cpu() * st["std"] + st["mean"], bins)


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-3), ("simt", 1e-4)])
def test_inference_against_the_reference(golden, mode, tol):
    ada = build(golden, mode)
    ref = golden["inference"]["out"]
    x, mask = golden["x"].to(DEV), golden["src_mask"].to(DEV)
    with torch.no_grad():
        r = ada(x.clone(), mask, {}, inference=True, noise=list(golden["inference"]["noise"]))
        flips = int((r["duration_rounded"].cpu() != ref["duration_rounded"]).sum())
        if flips:
            r = ada(x.clone(), mask, {}, inference=True, noise=list(golden["inference"]["noise"]),
                    force={"duration_rounded": ref["duration_rounded"]})
    assert flips <= 1
    assert torch.equal(r["duration_rounded"].cpu(), ref["duration_rounded"])
    assert torch.equal(r["tgt_mask"].cpu(), ref["tgt_mask"]) and r["x"].shape[1] % 64 == 0
    assert (r["duration_prediction"].cpu() - ref["duration_prediction"]).abs().max() < tol
    same = torch.ones(ref["tgt_mask"].shape, dtype=torch.bool)
    for i, v in enumerate(golden["cfg"]["variances"]):
        err = (r[f"variances_{v}"].cpu() - ref[f"variances_{v}"]).abs().max()
        print(f"fastdiff [{mode}] {v}: max err {float(err):.2e}")
        assert err < tol * (1 + 2 * i), v      # (later variances see the earlier ones' embeddings)
        if i + 1 < len(golden["cfg"]["variances"]) or True:
            same &= _buckets(golden, v, r[f"variances_{v}"]) == _buckets(golden, v, ref[f"variances_{v}"])
    assert same.float().mean() > 0.95
    # rows whose bucket decisions agree carry the same embeddings: x and out must match there
    assert (r["x"].cpu() - ref["x"])[same].abs().max() < tol
    assert (r["out"].cpu() - ref["out"])[same].abs().max() < tol
    assert r["duration_z"] is None and r["variances_pitch_z"] is None


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-3), ("simt", 1e-4)])
def test_teacher_forced_against_the_reference(golden, mode, tol):
    ada = build(golden, mode)
    tf = golden["teacher_forced"]
    steps = dict(zip(["duration"] + golden["cfg"]["variances"], tf["steps"]))
    with torch.no_grad():
        r = ada(golden["x"].to(DEV), golden["src_mask"].to(DEV), tf["targets"], inference=False, noise=list(tf["noise"]),
                steps=steps, jitter=tf["jitter"])
    ref = tf["out"]
    assert torch.equal(r["tgt_mask"].cpu(), ref["tgt_mask"])
    for k in ("duration_prediction", "variances_pitch", "variances_energy", "x", "out"):
        err = (r[k].cpu() - ref[k]).abs().max()
        print(f"fastdiff teacher-forced [{mode}] {k}: max err {float(err):.2e}")
        assert err < tol, k
    for k in ("duration_z", "variances_pitch_z", "variances_energy_z"):
        assert torch.equal(r[k].cpu(), ref[k])   # the injected draws come back as the regression targets


def test_length_regulator_pad_to_multiple_is_bit_exact(golden):
    x = golden["x"].to(DEV)
    dur = golden["inference"]["out"]["duration_rounded"].to(DEV)
    out, mask = LengthRegulator(pad_to_multiple_of=64)(x, dur, 2756.25)
    ref = golden["inference"]["out"]
    assert out.shape[1] % 64 == 0 and torch.equal(mask.cpu(), ref["tgt_mask"])
    # truncation: an utterance longer than int(max_length) keeps its frames up to the ROUNDED length (model.py:355-369)
    from oracle import fastdiff_oracle as FO

    g = torch.Generator().manual_seed(3)
    x2 = torch.randn(3, 9, 32, generator=g)
    d2 = torch.tensor([[30, 40, 50, 0, 0, 0, 0, 0, 0], [5, 5, 5, 5, 5, 5, 5, 5, 60], [1, 0, 0, 0, 0, 0, 0, 0, 0]])
    want, wmask = FO.length_regulator_padded(x2, d2, 100.5, 64)
    got, gmask = LengthRegulator(pad_to_multiple_of=64)(x2.to(DEV), d2.to(DEV), 100.5)
    assert got.shape == want.shape == (3, 128, 32)
    assert torch.equal(got.cpu(), want) and torch.equal(gmask.cpu(), wmask)


def test_fastspeech2_with_the_fastdiff_adaptor_end_to_end():
    """FastSpeech2(fastdiff_variances=True): forward(inference=True) runs the diffusion adaptor between encoder and
    decoder (reference fastspeech2.py:302-320, 769-776); same noise -> same result; the train step raises"""
    kw = dict(configs.C2, fastdiff_variances=True, variance_nlayers=[2, 2], duration_nlayers=2)
    st = {v: {"min": -3.0, "max": 3.0, "mean": 0.0, "std": 1.0} for v in kw["variances"]}
    model = FastSpeech2(stats=st, phone2id={f"p{i}": i for i in range(80)}, num_workers=0, **kw)
    model.load_state_dict(synthetic.fill_state_dict(model.state_dict(), seed=2))
    model = model.eval().to(DEV)
    batch = synthetic.make_batch(3, 6, 14, seed=2)
    torch.manual_seed(0)
    with torch.no_grad():
        r1 = model(batch, inference=True)
    torch.manual_seed(0)
    with torch.no_grad():
        r2 = model(batch, inference=True)
    assert r1["mel"].shape[1] % 64 == 0 and r1["mel"].shape[2] == 80 and torch.isfinite(r1["mel"][~r1["tgt_mask"]]).all()
    assert torch.equal(r1["mel"], r2["mel"]) and torch.equal(r1["duration_rounded"], r2["duration_rounded"])
    assert {"variances_pitch_z", "variances_energy_z", "duration_z"} <= set(r1) and r1["duration_z"] is None
    assert (r1["duration_rounded"][r1["src_mask"]] == 0).all()
    model.train()
    with pytest.raises(NotImplementedError):
        model(synthetic.add_train_targets(batch, kw["variances"], seed=2), inference=False)
