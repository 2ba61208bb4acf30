"""Stochastic duration predictor, inference direction, on the CUDA path against the golden of the unmodified
reference module (oracle/make_goldens_sdp.py) and the oracle, with the reference's noise draw injected."""
import os

import pytest
import torch

from lightningfastspeech2_b200 import configs, ops, synthetic
from lightningfastspeech2_b200.fastspeech2.fastspeech2 import FastSpeech2
from lightningfastspeech2_b200.fastspeech2.model import StochasticDurationPredictorWrapper
from oracle import sdp_oracle as S

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(__file__), "golden", "sdp_small.pt")


def _module(g):
    c = g["cfg"]
    mod = StochasticDurationPredictorWrapper(c["nlayers"], c["in_channels"], c["filter_size"], c["kernel_size"], c["dropout"])
    sd = synthetic.fill_state_dict(g["shapes"], seed=g["seed"])
    mod.load_state_dict(sd, strict=True)              # same key names and shapes as the reference module
    return mod.eval().to(DEV), sd


def test_log_durations_match_the_reference_golden():
    g = torch.load(GOLD, weights_only=False)
    mod, _ = _module(g)
    for case in g["cases"]:
        with torch.no_grad():
            logw = mod(case["x"].to(DEV), case["src_mask"].to(DEV), sigma=case["sigma"], inference=True,
                       noise=case["noise"].to(DEV))
        err = float((logw.cpu() - case["logw"]).abs().max())
        print(f"sdp B={case['x'].shape[0]} Tp={case['x'].shape[1]}: max |logw - reference| = {err:.2e}")
        assert err < 1e-4, err                          # fp32 CUDA-core GEMMs + fp32 kernels vs torch CPU fp32
        assert torch.equal(logw.cpu() == 0, case["logw"] == 0)
        dur = ops.sdp_durations(logw, case["src_mask"].to(DEV)).cpu()
        # ceil() is discontinuous: a duration may only differ where exp(logw) sits within the float error of an integer
        w = torch.exp(case["logw"].double())
        near = (w - torch.round(w)).abs() < 2e-4 * w.clamp_min(1.0)
        assert torch.equal(dur[~near], case["duration_rounded"][~near])
        # exact on the reference's own log-durations (isolates the rounding / guard kernel)
        assert torch.equal(ops.sdp_durations(case["logw"].to(DEV), case["src_mask"].to(DEV)).cpu(), case["duration_rounded"])


def test_zero_duration_guard_and_sampling_without_injected_noise():
    logw = torch.full((2, 6), -30.0)                  # exp -> 0 -> ceil(1e-13...) = 1?  exp(-30) > 0 so ceil gives 1
    logw[1] = 0.0                                     # logw == 0 -> 0 everywhere -> guard sets the valid ones to 1
    mask = torch.zeros(2, 6, dtype=torch.bool)
    mask[1, 4:] = True
    dur = ops.sdp_durations(logw.to(DEV), mask.to(DEV)).cpu()
    assert torch.equal(dur, S.stochastic_durations(logw, mask))
    g = torch.load(GOLD, weights_only=False)
    mod, _ = _module(g)
    case = g["cases"][0]
    with torch.no_grad():
        a = mod(case["x"].to(DEV), case["src_mask"].to(DEV), inference=True)
        b = mod(case["x"].to(DEV), case["src_mask"].to(DEV), inference=True)
    assert torch.isfinite(a).all() and not torch.equal(a, b)          # sampled on the device, a new draw per call
    assert bool((a[case["src_mask"].to(DEV)] == 0).all())
    with pytest.raises(NotImplementedError):
        mod(case["x"].to(DEV), case["src_mask"].to(DEV), tgt=torch.ones(2, 13, device=DEV))   # training direction


def test_fastspeech2_with_the_stochastic_duration_predictor_end_to_end():
    """duration_stochastic=True (reference fastspeech2.py:72, model.py:196-207): the model builds, loads a state_dict with
    the reference's key names, synthesises; the durations follow model.py:302-309 from the predicted log-durations, and
    the mel equals the pipeline driven with those durations forced (the predictor only decides durations)"""
    kw = dict(configs.PRESETS["C2"])
    kw.update(duration_stochastic=True, duration_depthwise_conv=False)
    hp = configs.resolve(kw)
    st = {v: {"min": -3.0, "max": 3.0, "mean": 0.0, "std": 1.0} for v in hp["variances"]}
    model = FastSpeech2(stats=st, phone2id={f"p{i}": i for i in range(80)}, num_workers=0, **kw)
    keys = [k for k in model.state_dict() if "duration_predictor.sdp." in k]
    assert any(k.endswith("sdp.flows.0.log_scale") for k in keys) and any("post_flows" in k for k in keys)
    sd = synthetic.fill_state_dict(model.state_dict(), seed=4)
    model.load_state_dict(sd)
    model = model.eval().to(DEV)
    batch = synthetic.make_batch(3, 10, 30, seed=12)
    b, tp = batch["phones"].shape
    noise = torch.randn(b, 2, tp, generator=torch.Generator().manual_seed(8))
    with torch.no_grad():
        r = model(batch, inference=True, force={"sdp_noise": noise.to(DEV), "sdp_sigma": 0.8})
    logw = r["duration_prediction"]
    assert bool((logw[r["src_mask"]] == 0).all()) and torch.isfinite(logw).all()
    assert torch.equal(r["duration_rounded"].cpu(), S.stochastic_durations(logw.cpu(), r["src_mask"].cpu()))
    assert r["mel"].shape[1] == int(r["duration_rounded"].sum(1).max()) and torch.isfinite(r["mel"][~r["tgt_mask"]]).all()
    with torch.no_grad():  # same durations forced through the call: identical mel (the predictor only decides durations)
        r2 = model(batch, inference=True, force={"duration_rounded": r["duration_rounded"], "sdp_noise": noise.to(DEV)})
    assert torch.equal(r2["mel"], r["mel"])
    with pytest.raises(NotImplementedError):
        model(synthetic.add_train_targets(batch, hp["variances"], seed=1), inference=False)
