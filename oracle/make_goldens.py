"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.pt by running the UNMODIFIED reference.

Run in the authoring container (needs /root/reference):
    python -m oracle.make_goldens
The reference has no golden vectors of its own, so these are outputs of the reference's
own modules (imported through oracle/ref_shim.py) on weights/inputs from
lightningfastspeech2_b200.synthetic (numpy PCG64, rebuilt anywhere from the seed).
Each file stores inputs' seeds/tensors, the parameter name->shape table, and the
reference outputs in fp32 (and from a .double() copy of the same model as ground truth).
"""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lightningfastspeech2_b200 import configs, synthetic  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
STATS = {"min": -2.5, "max": 3.5, "mean": 0.3, "std": 1.7}


def _ref_model(kwargs, seed, stats=None):
    f = ref_shim.import_fastspeech2()
    dsmod = sys.modules["litfass.dataset.datasets"]

    hp = configs.resolve(kwargs)

    class DS(ref_shim.FakeTTSDataset):
        def __init__(self, ds=None, **kw):
            super().__init__(ds, **kw)
            if stats is not None:
                self.stats = {v: dict(stats) for v in hp["variances"]}
            for p in hp["priors"]:
                self.stats[f"{p}_prior"] = dict(stats or {"min": -3.0, "max": 3.0})

    dsmod.TTSDataset = DS
    f.TTSDataset = DS
    model = ref_shim.build_reference(dict(kwargs, num_workers=0))
    sd = synthetic.fill_state_dict(model.state_dict(), seed=seed)
    model.load_state_dict(sd, strict=True)
    model.eval()
    return model, hp


def _bucket_idx(model, hp, result):
    out = {}
    for var in hp["variances"]:
        enc = model.variance_adaptor.encoders[var]
        out[var] = torch.bucketize(result[f"variances_{var}"] * enc.std + enc.mean, enc.bins)
    return out


def _run(model, batch, inference):
    with torch.no_grad():
        r32 = model(batch, inference=inference)
    m64 = copy.deepcopy(model).double()
    b64 = {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in batch.items()}
    if inference:
        # force the fp32 model's discrete decisions on the fp64 copy by feeding them as targets
        b64 = dict(b64)
    with torch.no_grad():
        r64 = m64(b64, inference=inference)
    return r32, r64


def grad_probe(name, n):
    """seeded +-1 probe vector for gradient fingerprints (rebuilt identically by the tests)"""
    import zlib

    r = np.random.default_rng([97, zlib.crc32(name.encode())])
    return torch.from_numpy(r.integers(0, 2, size=n).astype(np.float64) * 2 - 1)


def forward_golden(name, preset, seed, batch, inference, stats=None, train_targets=False):
    model, hp = _ref_model(configs.PRESETS[preset], seed, stats)
    if train_targets:
        batch = synthetic.add_train_targets(batch, hp["variances"], seed=seed, levels=hp["variance_levels"])
    if hp["priors"]:  # per-utterance scalar priors (the dataset emits python lists; tensors work the same way)
        batch = dict(batch)
        g = np.random.default_rng(seed + 1000)
        for p in hp["priors"]:
            batch[f"priors_{p}"] = torch.from_numpy(g.uniform(-2.0, 3.0, size=batch["phones"].shape[0]).astype(np.float32))
    r32, r64 = _run(model, batch, inference)
    g = {
        "preset": preset, "seed": seed, "inference": inference, "stats": stats,
        "shapes": {k: tuple(v.shape) for k, v in model.state_dict().items()},
        "batch": {k: v for k, v in batch.items() if torch.is_tensor(v)},
        "out": {k: v.clone() for k, v in r32.items() if torch.is_tensor(v)},
        "out64_mel": r64["mel"].clone(),
        "out64_duration_prediction": r64["duration_prediction"].clone(),
        "bucket_idx": _bucket_idx(model, hp, r32) if inference else None,
    }
    if inference and not torch.equal(r32["duration_rounded"], r64["duration_rounded"]):
        # keep the fp64 ground truth comparable: rerun fp64 teacher-forced on fp32's durations
        g["out64_mel"] = None
    if not inference:
        losses = model.loss(r32, batch)
        g["loss"] = {k: float(v) for k, v in losses.items()}
        model.zero_grad()
        model.train()
        for mod in model.modules():  # dropout off, as in the parity protocol (SURVEY 8c iv)
            if isinstance(mod, torch.nn.Dropout):
                mod.p = 0.0
            if isinstance(mod, torch.nn.MultiheadAttention):
                mod.dropout = 0.0
        r = model(batch, inference=False)
        ls = model.loss(r, batch)
        ls["total"].backward()
        g["loss_train_mode"] = {k: float(v) for k, v in ls.items()}
        g["grad_norms"] = {k: float(p.grad.norm()) for k, p in model.named_parameters() if p.grad is not None}
        # gradient fingerprints: the dot product with a seeded probe vector pins direction as well as
        # magnitude without committing megabytes; small tensors are stored whole
        g["grad_dots"] = {k: float((p.grad.double().flatten() * grad_probe(k, p.numel())).sum())
                          for k, p in model.named_parameters() if p.grad is not None}
        g["grad_small"] = {k: p.grad.clone() for k, p in model.named_parameters()
                           if p.grad is not None and p.numel() <= 512}
        # one optimizer step exactly as the reference configures it (fastspeech2.py:1166-1182)
        before = {k: p.detach().clone() for k, p in model.named_parameters()}
        (opt,), (sch,) = model.configure_optimizers()
        opt.step()
        sch["scheduler"].step()
        g["lr_first_step"] = float(opt.param_groups[0]["lr"])
        g["step_delta_norms"] = {k: float((p.detach() - before[k]).norm()) for k, p in model.named_parameters()}
        g["step_delta_dots"] = {k: float(((p.detach() - before[k]).double().flatten() * grad_probe(k, p.numel())).sum())
                                for k, p in model.named_parameters()}
    torch.save(g, os.path.join(OUT, name + ".pt"))
    valid = int((~r32["tgt_mask"]).sum())
    print(f"{name}: mel {tuple(r32['mel'].shape)} valid_frames {valid}")


def length_regulator_golden():
    m = ref_shim.import_model()
    lr = m.LengthRegulator()
    g = np.random.default_rng(7)
    cases = []

    def add(tag, x, dur, max_length):
        out, mask = lr(x, dur, max_length)
        cases.append({"tag": tag, "x": x, "dur": dur, "max_length": max_length, "out": out.clone(), "mask": mask.clone()})

    for i, (b, tp, d, hi, dt) in enumerate([(3, 7, 4, 5, torch.int64), (4, 16, 8, 10, torch.int32),
                                             (2, 1, 4, 3, torch.int64), (5, 33, 4, 4, torch.int32)]):
        x = torch.from_numpy(g.standard_normal((b, tp, d)).astype(np.float32))
        dur = torch.from_numpy(g.integers(0, hi + 1, size=(b, tp))).to(dt)
        dur[0, 0] = max(int(dur[0, 0]), 1)
        add(f"rand{i}", x, dur, 2756.25)
    x = torch.from_numpy(g.standard_normal((3, 6, 4)).astype(np.float32))
    dur = torch.tensor([[2, 0, 3, 1, 0, 0], [0, 0, 0, 0, 0, 0], [1, 1, 1, 1, 1, 1]])
    add("zero_row", x, dur, 2756.25)
    dur = torch.tensor([[2, 0, 30, 1, 0, 0], [1, 2, 3, 0, 0, 0], [5, 5, 5, 5, 5, 5]], dtype=torch.int32)
    add("truncate", x, dur, 12.75)
    xb = torch.from_numpy(g.standard_normal((3, 6, 8)).astype(np.float32)).to(torch.bfloat16)
    add("bf16", xb, dur, 20.0)
    xn = x.clone()
    xn[0, 1] = float("nan")  # a zero-duration phone holding NaN must not leak
    xn[1, 2] = -0.0
    add("nan_negzero", xn, torch.tensor([[2, 0, 3, 1, 0, 0], [1, 2, 3, 0, 0, 0], [1, 0, 0, 0, 0, 1]]), 2756.25)
    torch.save(cases, os.path.join(OUT, "length_regulator.pt"))
    print(f"length_regulator: {len(cases)} cases")


def block_golden():
    """One FFTBlock and one VariancePredictor of the reference on a padded batch (tiny + d=256)."""
    m = ref_shim.import_model()
    g = np.random.default_rng(11)
    out = []
    for tag, d, fsz, k, dw in [("dw_d32", 32, 64, 5, True), ("dense_d32", 32, 64, 3, False),
                               ("dw_d256", 256, 1024, 25, True)]:
        layer = m.ConformerEncoderLayer(d, 2, conv_in=d, conv_filter_size=fsz, conv_kernel=(k, 1),
                                        batch_first=True, dropout=0.1, conv_depthwise=dw).eval()
        sd = synthetic.fill_state_dict(layer.state_dict(), seed=3)
        layer.load_state_dict(sd)
        b, t = 2, 40
        x = torch.from_numpy(g.standard_normal((b, t, d)).astype(np.float32))
        kpm = torch.zeros(b, t, dtype=torch.bool)
        kpm[1, 29:] = True
        with torch.no_grad():
            y = layer(x, src_key_padding_mask=kpm)
            y64 = copy.deepcopy(layer).double()(x.double(), src_key_padding_mask=kpm)
        out.append({"tag": tag, "d": d, "filter": fsz, "k": k, "depthwise": dw, "x": x, "kpm": kpm,
                    "shapes": {n: tuple(v.shape) for n, v in sd.items()}, "y": y, "y64": y64})
    torch.save(out, os.path.join(OUT, "fft_block.pt"))
    print("fft_block:", [o["tag"] for o in out])


def variance_encoder_golden():
    """The reference's VarianceEncoder called directly, with and without a target and with `control` != 1 (the
    adaptor never passes it, model.py:284-288/323-327, so the whole-model goldens cannot pin that argument)."""
    m = ref_shim.import_model()
    g = np.random.default_rng(21)
    out = []
    for tag, d, nl, dw in [("dw_d32", 32, 2, True), ("dense_d32", 32, 2, False), ("dw_d256", 256, 5, True)]:
        enc = m.VarianceEncoder(nl, d, d, 3, 0.1, dw, STATS["min"], STATS["max"], STATS["mean"], STATS["std"], 256,
                                False).eval()
        sd = synthetic.fill_state_dict(enc.state_dict(), seed=5)
        enc.load_state_dict(sd)
        b, t = 2, 37
        x = torch.from_numpy(g.standard_normal((b, t, d)).astype(np.float32))
        tgt = torch.from_numpy(g.standard_normal((b, t)).astype(np.float32))
        mask = torch.zeros(b, t, dtype=torch.bool)
        mask[1, 25:] = True
        with torch.no_grad():
            p_tf, e_tf = enc(x, tgt, mask)
            p_free, e_free = enc(x, None, mask)
            p_ctl, e_ctl = enc(x, None, mask, control=1.3)
        out.append({"tag": tag, "d": d, "nlayers": nl, "depthwise": dw, "x": x, "tgt": tgt, "mask": mask,
                    "mean": STATS["mean"], "std": STATS["std"], "shapes": {n: tuple(v.shape) for n, v in sd.items()},
                    "teacher_forced": (p_tf, e_tf), "free": (p_free, e_free), "control_1.3": (p_ctl, e_ctl)})
    torch.save(out, os.path.join(OUT, "variance_encoder.pt"))
    print("variance_encoder:", [o["tag"] for o in out])


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    length_regulator_golden()
    block_golden()
    variance_encoder_golden()
    tiny = synthetic.make_batch(3, 5, 17, seed=5)
    forward_golden("tiny_dw_infer", "TINY_DW", 1, tiny, True, stats=STATS)
    forward_golden("tiny_dense_infer", "TINY_DENSE", 2, tiny, True, stats=STATS)
    forward_golden("tiny_dw_train", "TINY_DW", 3, tiny, False, stats=STATS, train_targets=True)
    forward_golden("small_train", "SMALL_TRAIN", 4, synthetic.make_batch(3, 9, 40, seed=4), False, stats=STATS,
                   train_targets=True)
    forward_golden("small_train_phone", "SMALL_TRAIN_PHONE", 6, synthetic.make_batch(3, 9, 40, seed=6), False, stats=STATS,
                   train_targets=True)
    forward_golden("small_train_dense", "SMALL_TRAIN_DENSE", 8, synthetic.make_batch(3, 9, 40, seed=8), False, stats=STATS,
                   train_targets=True)
    forward_golden("small_train_prior", "SMALL_TRAIN_PRIOR", 9, synthetic.make_batch(3, 9, 40, seed=9), False, stats=STATS,
                   train_targets=True)
    forward_golden("small_phone_infer", "SMALL_TRAIN_PHONE", 7, synthetic.make_batch(3, 9, 40, seed=7), True, stats=STATS)
    forward_golden("c1_infer", "C1", 1234, synthetic.make_batch(1, 128, 128, seed=1234), True)
    forward_golden("c2_small_infer", "C2", 2, synthetic.make_batch(4, 20, 96, seed=2), True)


if __name__ == "__main__":
    main()
