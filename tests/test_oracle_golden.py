"""Pin oracle/fs2_oracle.py against outputs of the reference itself (tests/golden/*.pt,
written by oracle/make_goldens.py from the unmodified reference under oracle/ref_shim.py).

The reference ships no tests/golden vectors, so these are the pin."""
import os

import pytest
import torch

from lightningfastspeech2_b200 import configs, synthetic
from oracle import fs2_oracle as O

FWD = ["tiny_dw_infer", "tiny_dense_infer", "c1_infer", "c2_small_infer", "small_phone_infer"]


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False)


def _setup(g):
    hp = configs.resolve(configs.PRESETS[g["preset"]])
    if g["stats"] is not None:
        hp["stats"] = {v: g["stats"] for v in hp["variances"]}
    sd = synthetic.fill_state_dict(g["shapes"], seed=g["seed"], stats=g["stats"])
    return hp, sd


@pytest.mark.parametrize("name", FWD)
def test_forward_inference_fp32(golden_dir, name):
    g = _load(golden_dir, name)
    hp, sd = _setup(g)
    ref = g["out"]
    # free-running: discrete decisions must agree with the reference's own fp32 run
    r = O.forward(sd, hp, g["batch"], inference=True)
    assert r["duration_rounded"].dtype == torch.int32
    assert torch.equal(r["duration_rounded"], ref["duration_rounded"])
    assert torch.equal(r["tgt_mask"], ref["tgt_mask"])
    assert torch.equal(r["src_mask"], ref["src_mask"])
    for var in hp["variances"]:
        flips = (r[f"_bucket_{var}"] != g["bucket_idx"][var]).sum().item()
        assert flips == 0, f"{var}: {flips} bucket flips"
    # tolerance: both sides are fp32 CPU torch; only summation order differs
    assert (r["mel"] - ref["mel"]).abs().max() < 2e-5
    assert (r["duration_prediction"] - ref["duration_prediction"]).abs().max() < 1e-5
    for var in hp["variances"]:
        assert (r[f"variances_{var}"] - ref[f"variances_{var}"]).abs().max() < 1e-5
    assert (r["fastdiff_var"] - ref["fastdiff_var"]).abs().max() < 1e-5


@pytest.mark.parametrize("name", FWD)
def test_forward_inference_fp64_truth(golden_dir, name):
    """fp64 oracle (discrete decisions forced to the reference's) vs the reference's own
    .double() forward: restatement is exact up to fp64 rounding."""
    g = _load(golden_dir, name)
    if g["out64_mel"] is None:
        pytest.skip("fp64 reference made different discrete decisions")
    hp, sd = _setup(g)
    b = dict(g["batch"])
    b["speaker"] = b["speaker"].double()
    r = O.forward(sd, hp, b, inference=True, dtype=torch.float64)
    if not torch.equal(r["tgt_mask"], g["out"]["tgt_mask"]):
        pytest.skip("fp64 durations differ from fp32 reference")
    valid = ~g["out"]["tgt_mask"]
    err = (r["mel"] - g["out64_mel"]).abs()
    # bucket flips between fp32/fp64 would show as O(1) errors; bound on frames without flips
    assert err[valid].max() < 1e-9 or err[valid].median() < 1e-12
    assert (r["duration_prediction"] - g["out64_duration_prediction"]).abs().max() < 1e-10


def test_forward_train_and_loss(golden_dir):
    g = _load(golden_dir, "tiny_dw_train")
    hp, sd = _setup(g)
    r = O.forward(sd, hp, g["batch"], inference=False)
    assert torch.equal(r["duration_rounded"], g["batch"]["duration"])
    assert (r["mel"] - g["out"]["mel"]).abs().max() < 2e-5
    ls = O.loss(hp, r, g["batch"])
    for k, v in g["loss"].items():
        assert abs(float(ls[k]) - v) <= 1e-5 * max(1.0, abs(v)), k


def test_fft_block(golden_dir):
    for c in _load(golden_dir, "fft_block"):
        sd = synthetic.fill_state_dict(c["shapes"], seed=3)
        y = O.fft_block(c["x"], c["kpm"], sd, "", 2, c["depthwise"], torch.float32)
        assert (y - c["y"]).abs().max() < 1e-5, c["tag"]
        sd64 = {k: v.double() for k, v in sd.items()}
        y64 = O.fft_block(c["x"].double(), c["kpm"], sd64, "", 2, c["depthwise"], torch.float64)
        assert (y64 - c["y64"]).abs().max() < 1e-12, c["tag"]


def test_variance_encoder_with_control(golden_dir):
    """VarianceEncoder called directly (model.py:409-441): teacher-forced, free-running and with control = 1.3 (the
    embedding follows the UNSCALED prediction, only the returned prediction is scaled)"""
    for c in _load(golden_dir, "variance_encoder"):
        sd = synthetic.fill_state_dict(c["shapes"], seed=5)
        for key, tgt, ctl in (("teacher_forced", c["tgt"], 1.0), ("free", None, 1.0), ("control_1.3", None, 1.3)):
            pred, emb, _ = O.variance_encoder(c["x"], tgt, c["mask"], sd, "", c["nlayers"], c["depthwise"], c["mean"],
                                              c["std"], torch.float32, control=ctl)
            rp, re = c[key]
            assert (pred - rp).abs().max() < 1e-5, (c["tag"], key)
            assert torch.equal(emb, re), (c["tag"], key)
        assert torch.equal(c["control_1.3"][1], c["free"][1])  # same buckets with and without control


def test_length_regulator_bit_exact(golden_dir):
    for c in _load(golden_dir, "length_regulator"):
        out, mask = O.length_regulator(c["x"], c["dur"], c["max_length"])
        assert out.dtype == c["out"].dtype and out.shape == c["out"].shape, c["tag"]
        assert torch.equal(mask, c["mask"]), c["tag"]
        # bit-exact incl. NaN payloads and the sign of zero
        a = out.contiguous().view(torch.int16 if out.dtype == torch.bfloat16 else torch.int32)
        b = c["out"].contiguous().view(torch.int16 if out.dtype == torch.bfloat16 else torch.int32)
        assert torch.equal(a, b), c["tag"]


def test_noam_matches_formula():
    assert abs(O.noam_scale(0, 4000) - O.noam_scale(1, 4000)) == 0
    assert abs(O.noam_scale(4000, 4000) - 1.0) < 1e-12
    assert O.noam_scale(16000, 4000) == pytest.approx(0.5)


def _probe(name, n):
    import zlib

    import numpy as np

    r = np.random.default_rng([97, zlib.crc32(name.encode())])
    return torch.from_numpy(r.integers(0, 2, size=n).astype(np.float64) * 2 - 1)


@pytest.mark.parametrize("name", ["tiny_dw_train", "small_train", "small_train_phone", "small_train_dense", "small_train_prior"])
def test_train_step_gradients_and_adamw(golden_dir, name):
    """Oracle autograd vs the reference's own backward() + AdamW/Noam step: loss values, per-parameter
    gradient norms, probe-vector dot products (direction), small tensors whole, and the first
    optimizer step's parameter deltas."""
    g = _load(golden_dir, name)
    hp, sd = _setup(g)
    losses, grads = O.gradients(sd, hp, g["batch"])
    for k, v in g["loss_train_mode"].items():
        assert abs(losses[k] - v) <= 1e-5 * max(1.0, abs(v)), k
    assert set(g["grad_norms"]) - {k for k in g["grad_norms"] if k.startswith("fastdiff_linear")} <= set(grads)
    scale = max(g["grad_norms"].values())
    for k, ref in g["grad_norms"].items():
        if k.startswith("fastdiff_linear"):
            continue
        gk = grads[k]
        assert abs(float(gk.norm()) - ref) <= 2e-4 * max(ref, 1e-3 * scale), k
        dot = float((gk.double().flatten() * _probe(k, gk.numel())).sum())
        assert abs(dot - g["grad_dots"][k]) <= 2e-4 * max(ref * gk.numel() ** 0.5, 1e-3 * scale), k
    for k, ref in g["grad_small"].items():
        if k in grads:
            assert (grads[k] - ref).abs().max() <= 1e-4 * max(float(ref.abs().max()), 1e-3 * scale), k
    # first AdamW step with the Noam lr (scheduler not yet stepped)
    for k, ref in g["step_delta_norms"].items():
        if k not in grads:
            continue
        p0 = sd[k]
        p1, _, _, lr = O.adamw_noam_step(p0, grads[k], torch.zeros_like(p0), torch.zeros_like(p0), 1, hp["lr"],
                                         hp["warmup_steps"])
        delta = p1 - p0
        assert abs(float(delta.norm()) - ref) <= 2e-3 * max(ref, 1e-12), k
    assert abs(O.noam_scale(1, hp["warmup_steps"]) * hp["lr"] - g["lr_first_step"]) < 1e-15
