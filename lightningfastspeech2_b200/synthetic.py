"""Seeded synthetic weights and utterance batches (SURVEY.md 8d "Synthetic inputs").

There is no network for checkpoints or LibriTTS, so benches, tests and the golden
vectors all use weights/inputs generated here.  Everything is drawn from numpy PCG64
streams keyed on (seed, crc32(parameter name)), so a tensor depends only on its name,
shape and the seed -- the same state_dict can be rebuilt anywhere without the reference.
"""
import math
import zlib

import numpy as np
import torch


def _rng(seed, name):
    return np.random.default_rng([int(seed) & 0xFFFFFFFF, zlib.crc32(name.encode())])


def sinusoid_table(max_len, d):
    """PositionalEncoding table (reference model.py:43-51): interleaved sin/cos."""
    pe = torch.zeros(max_len, d)
    pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2).float() * (-math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.unsqueeze(0)


def fill_state_dict(template, seed=0, duration_bias=math.log(6.0), duration_scale=0.25, stats=None):
    """Return {name: tensor} with the shapes of ``template`` -- a state_dict, or a
    {name: shape tuple} table (then ``bins`` are rebuilt as linspace(min, max, nbins-1)
    from ``stats`` [default min -3 / max 3] and ``pe`` from the sinusoid formula).

    Init mimics the scale of torch's defaults (so activations look like a freshly
    constructed reference model) but perturbs LayerNorm affine terms and biases away
    from 1/0 so that parity tests exercise them.  ``bins`` and ``pe`` are structural
    and are copied from the template.  The duration head is calibrated as SURVEY 8d
    prescribes (weight *= 0.25, bias = log 6) so predicted durations are ~2..9 frames.
    """
    out = {}
    for name, ref in template.items():
        if not torch.is_tensor(ref):
            shape = tuple(ref)
            if name.endswith(".bins"):
                st = (stats or {"min": -3.0, "max": 3.0})
                out[name] = torch.linspace(st["min"], st["max"], shape[0])
                continue
            if name.endswith("positional_encoding.pe"):
                out[name] = sinusoid_table(shape[1], shape[2])
                continue
            ref = torch.empty(0, dtype=torch.float32)
        else:
            shape = tuple(ref.shape)
            if name.endswith(".bins") or name.endswith("positional_encoding.pe"):
                out[name] = ref.detach().clone()
                continue
        g = _rng(seed, name)
        leaf = name.rsplit(".", 1)[-1]
        is_ln = (".norm1." in name or ".norm2." in name or name.endswith("layers.2.weight")
                 or name.endswith("layers.2.bias"))
        if is_ln and leaf == "weight":
            a = 1.0 + 0.1 * g.standard_normal(shape)
        elif is_ln and leaf == "bias":
            a = 0.1 * g.standard_normal(shape)
        elif leaf == "gamma":                       # LayerNorm2 of the stochastic duration predictor
            a = 1.0 + 0.1 * g.standard_normal(shape)
        elif leaf in ("beta", "translation", "log_scale"):
            a = 0.1 * g.standard_normal(shape)
        elif "embedding.weight" in name and "speaker_embedding" not in name:
            a = g.standard_normal(shape)
            if name == "phone_embedding.weight":
                a[0] = 0.0  # padding_idx row
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            bound = 1.0 / math.sqrt(fan_in)
            a = g.uniform(-bound, bound, shape)
        else:
            a = g.uniform(-0.05, 0.05, shape)
        t = torch.from_numpy(np.asarray(a, dtype=np.float64)).to(ref.dtype)
        out[name] = t
    k = "variance_adaptor.duration_predictor.linear."
    if k + "weight" in out:
        out[k + "weight"] = out[k + "weight"] * duration_scale
        out[k + "bias"] = torch.full_like(out[k + "bias"], duration_bias)
    return out


def make_batch(batch_size, min_len, max_len, seed=0, vocab=80, pad_to=None):
    """Synthesis batch in the layout TTSDataset._collate_fn emits
    (reference dataset/datasets.py:852-882): ``phones`` (B,Tp) int64 with 0 = PAD,
    ``speaker`` (B,256) float32 d-vectors."""
    g = _rng(seed, "batch")
    lens = g.integers(min_len, max_len + 1, size=batch_size)
    tp = int(lens.max()) if pad_to is None else int(pad_to)
    phones = np.zeros((batch_size, tp), dtype=np.int64)
    for b, n in enumerate(lens):
        phones[b, :n] = g.integers(1, vocab, size=n)
    speaker = g.standard_normal((batch_size, 256)).astype(np.float32)
    return {
        "phones": torch.from_numpy(phones),
        "speaker": torch.from_numpy(speaker),
        "phones_lengths": torch.from_numpy(lens.astype(np.int64)),
    }


def add_train_targets(batch, variances, seed=0, n_mels=80, dur_lo=1, dur_hi=9, levels=None):
    """Teacher-forcing targets for the train-step config:
    duration ~ U{dur_lo..dur_hi} on valid phones, Tm = max sum(duration) exactly
    (HEAD quirk 6), mel ~ N(0,1), variances_* ~ N(0,1)."""
    g = _rng(seed, "targets")
    phones = batch["phones"].numpy()
    valid = phones != 0
    dur = g.integers(dur_lo, dur_hi + 1, size=phones.shape) * valid
    tm = int(dur.sum(1).max())
    b = phones.shape[0]
    out = dict(batch)
    out["duration"] = torch.from_numpy(dur.astype(np.int64))
    out["mel"] = torch.from_numpy(g.standard_normal((b, tm, n_mels)).astype(np.float32))
    for i, v in enumerate(variances):
        frame = levels is None or levels[i] == "frame"
        shape = (b, tm) if frame else phones.shape  # phone-level targets are per phoneme (datasets.py:592-601)
        out[f"variances_{v}"] = torch.from_numpy(g.standard_normal(shape).astype(np.float32))
    return out


def hifigan_state_dict(cfg, seed=0):
    """Seeded weights of the HiFi-GAN generator (reference third_party/hifigan/models.py:112-148) with weight_norm
    folded: {"conv_pre.weight", "ups.i.weight" (C_in, C_out, K), "resblocks.j.convs{1,2}.m.weight", "conv_post.weight",
    + biases}.  Gains are chosen like the bundled universal checkpoint's (std * sqrt(fan_in) of
    order 1) so that the signal survives the 4 x 9 residual convolutions and tanh is exercised but not saturated
    (|wav| mean ~0.3, max ~0.9 on N(0,1) mels)."""
    out = {}

    def conv(name, shape, fan_in, gain):
        g = _rng(seed, name)
        out[name + ".weight"] = torch.from_numpy((g.standard_normal(shape) * gain / math.sqrt(fan_in)).astype(np.float32))
        nb = shape[1] if name.startswith("ups.") else shape[0]
        out[name + ".bias"] = torch.from_numpy((g.standard_normal(nb) * 0.05).astype(np.float32))

    ch = cfg["upsample_initial_channel"]
    conv("conv_pre", (ch, cfg["num_mels"], 7), cfg["num_mels"] * 7, 1.0)
    nk = len(cfg["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        c_in, c_out = ch // (2 ** i), ch // (2 ** (i + 1))
        conv(f"ups.{i}", (c_in, c_out, k), c_in * k // u, 1.2)
        for j, (ks, ds) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            for m in range(len(ds)):
                conv(f"resblocks.{i * nk + j}.convs1.{m}", (c_out, c_out, ks), c_out * ks, 0.8)
                conv(f"resblocks.{i * nk + j}.convs2.{m}", (c_out, c_out, ks), c_out * ks, 0.6)
    conv("conv_post", (1, ch // (2 ** len(cfg["upsample_rates"])), 7), ch // (2 ** len(cfg["upsample_rates"])) * 7, 1.0)
    return out
