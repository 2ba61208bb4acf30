#!/usr/bin/env python
"""Operand-rounding emulation of the tensor-core precision recipes (CPU, fp64).

The teacher-forced mel path is evaluated in fp64 through the oracle with ONLY the operands of chosen products
rounded the way a tensor-core recipe would round them (accumulation stays fp64: the fp32 accumulator's own error,
~1e-6, is measured separately by the `simt` mode on the GPU).  The result is compared with the un-rounded fp64 run:
max |d mel| over all positions.  Products: `gemm` = every Linear / pointwise conv (QKV, out-proj, FFN, predictors,
mel), `qk` = Q.K^T, `pv` = P.V.  Recipes per operand pair (A = activation / Q / P, B = weight / K / V):

    x3      bf16 hi/lo split of both, hi.hi + lo.hi + hi.lo          (3 MMA passes; today's fp32-parity mode)
    a16b    A as ONE fp16 value, B as bf16 hi/lo: a.hi + a.lo         (2 passes)
    abf_b   A as ONE bf16 value, B as bf16 hi/lo                      (2 passes)
    a_bbf   A as bf16 hi/lo, B as ONE bf16 value                      (2 passes)
    a_b16   A as bf16 hi/lo, B as ONE fp16 value                      (2 passes, needs fp16 B planes)
    a16w16  A as ONE fp16 value, B as fp16 hi/lo: a.hi + a.lo         (2 passes; what npass = 2 of lfs2_gemm_tc_ex /
            lfs2_ffn_fused_tc_ex runs -- one MMA cannot mix an fp16 A with a bf16 B, so the weights are fp16 pairs)
    x1      both as one bf16 value                                    (1 pass; bf16 mode)
    x1h     both as one fp16 value                                    (1 pass)

    python tools/precision_emulation.py [preset] > profiles/<round>_precision_emulation.txt
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from lightningfastspeech2_b200 import configs, synthetic  # noqa: E402
from oracle import fs2_oracle as O  # noqa: E402


def bf(x):
    return x.float().bfloat16().double()


def f16(x):
    return x.float().half().double()


def split(x):
    hi = bf(x)
    return hi, bf(x - hi)


def product(a, b, recipe, mm):
    """mm(a, b) with the operands rounded per `recipe` (a, b fp64)"""
    if recipe == "exact":
        return mm(a, b)
    if recipe == "x3":
        ah, al = split(a)
        bh, bl = split(b)
        return mm(ah, bh) + mm(al, bh) + mm(ah, bl)
    if recipe == "a16b":
        bh, bl = split(b)
        a1 = f16(a)
        return mm(a1, bh) + mm(a1, bl)
    if recipe == "a16w16":
        bh = f16(b)
        bl = f16(b - bh)
        a1 = f16(a)
        return mm(a1, bh) + mm(a1, bl)
    if recipe == "abf_b":
        bh, bl = split(b)
        a1 = bf(a)
        return mm(a1, bh) + mm(a1, bl)
    if recipe == "a_bbf":
        ah, al = split(a)
        b1 = bf(b)
        return mm(ah, b1) + mm(al, b1)
    if recipe == "a_b16":
        ah, al = split(a)
        b1 = f16(b)
        return mm(ah, b1) + mm(al, b1)
    if recipe == "x1":
        return mm(bf(a), bf(b))
    if recipe == "x1h":
        return mm(f16(a), f16(b))
    raise ValueError(recipe)


class Emu:
    def __init__(self, gemm="exact", qk="exact", pv="exact", sites=None):
        """sites: optional {site: recipe} overriding `gemm` at one GEMM site: "qkv" (attention in-projection), "out"
        (attention out-projection), "ffn1" (the FFN's widening convolution, k>1 or d -> d_hidden), "ffn2" (the FFN's second, k=1 convolution)"""
        self.r = {"gemm": gemm, "qk": qk, "pv": pv}
        self.sites = sites or {}

    def __enter__(self):
        self.saved = (F.linear, F.conv1d, O.self_attention)
        lin, conv, r, sites = F.linear, F.conv1d, self.r, self.sites

        def linear(x, w, b=None, site=None):
            if x.dtype != torch.float64 or w.shape[0] == 1:       # (the predictor heads are CUDA-core row dots)
                return lin(x, w, b)
            y = product(x, w, sites.get(site, r["gemm"]), lambda a, bb: lin(a, bb))
            return y if b is None else y + b

        def conv1d(x, w, b=None, stride=1, padding=0, dilation=1, groups=1):
            if x.dtype != torch.float64 or groups != 1:           # depthwise / grouped 1x1: CUDA cores (fp32)
                return conv(x, w, b, stride, padding, dilation, groups)
            site = "ffn1" if (w.shape[2] > 1 or w.shape[0] > w.shape[1]) else ("ffn2" if w.shape[1] > w.shape[0] else None)
            y = product(x, w, sites.get(site, r["gemm"]), lambda a, bb: conv(a, bb, None, stride, padding, dilation, groups))
            return y if b is None else y + b[None, :, None]

        def self_attention(x, kpm, w_in, b_in, w_out, b_out, nhead):
            bsz, t, d = x.shape
            dh = d // nhead
            qkv = linear(x, w_in, b_in, site="qkv")
            q, k, v = qkv.split(d, dim=-1)
            heads = lambda z: z.reshape(bsz, t, nhead, dh).permute(0, 2, 1, 3)
            q, k, v = heads(q), heads(k), heads(v)
            s = product(q, k, r["qk"], lambda a, bb: torch.matmul(a, bb.transpose(-1, -2))) * (dh ** -0.5)
            if kpm is not None:
                s = s.masked_fill(kpm[:, None, None, :], float("-inf"))
            m = s.amax(-1, keepdim=True)
            p = torch.exp(s - m)                                  # un-normalised, as the flash kernel holds it
            l = p.sum(-1, keepdim=True)                           # row sum of the UNROUNDED probabilities (fp32 in the kernel)
            a = product(p, v, r["pv"], lambda a_, bb: torch.matmul(a_, bb)) / l
            return linear(a.permute(0, 2, 1, 3).reshape(bsz, t, d), w_out, b_out, site="out")

        F.linear, F.conv1d, O.self_attention = linear, conv1d, self_attention
        return self

    def __exit__(self, *exc):
        F.linear, F.conv1d, O.self_attention = self.saved


def main():
    preset = sys.argv[1] if len(sys.argv) > 1 else "C2"
    kw = configs.PRESETS[preset]
    hp = configs.resolve(kw)
    hp["stats"] = {v: {"min": -3.0, "max": 3.0, "mean": 0.0, "std": 1.0} for v in hp["variances"]}
    from lightningfastspeech2_b200.fastspeech2.fastspeech2 import FastSpeech2

    model = FastSpeech2(stats=hp["stats"], phone2id={f"p{i}": i for i in range(80)}, num_workers=0, **kw)
    sd = synthetic.fill_state_dict(model.state_dict(), seed=0)
    bsz, lo, hi = (4, 40, 96) if preset != "C3" else (2, 30, 60)
    batch = synthetic.add_train_targets(synthetic.make_batch(bsz, lo, hi, seed=7), hp["variances"], seed=7)
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        ref = O.forward(sd, hp, batch, inference=False, dtype=torch.float64)["mel"]
        f32 = O.forward(sd, hp, batch, inference=False, dtype=torch.float32)["mel"]
    print(f"# {preset}: teacher-forced mel path, B={bsz}, phones {lo}..{hi}, mel {tuple(ref.shape)}, |mel|max {float(ref.abs().max()):.2f}")
    print(f"# reference fp32 forward vs fp64: {float((f32.double() - ref).abs().max()):.2e}")
    print("| gemm | Q.K^T | P.V | MMA passes (gemm / qk / pv) | max abs mel error vs fp64 |")
    print("|---|---|---|---|---|")
    passes = {"exact": "-", "x3": 3, "a16b": 2, "abf_b": 2, "a_bbf": 2, "a_b16": 2, "x1": 1, "x1h": 1}
    rows = [("x3", "x3", "x3"), ("x3", "x3", "a16b"), ("x3", "x3", "abf_b"), ("x3", "x3", "x1"), ("x3", "x3", "x1h"),
            ("x3", "a16b", "a16b"), ("x3", "a_b16", "a16b"), ("x3", "x1h", "a16b"), ("x3", "x1h", "x1h"), ("x3", "x1", "x1"),
            ("a_b16", "x3", "x3"), ("a16b", "x3", "x3"), ("a_bbf", "x3", "x3"), ("abf_b", "x3", "x3"),
            ("a_b16", "x3", "a16b"), ("x1h", "x1h", "x1h"), ("x1", "x1", "x1")]
    for gm, qk, pv in rows:
        with Emu(gm, qk, pv), torch.no_grad():
            out = O.forward(sd, hp, batch, inference=False, dtype=torch.float64)["mel"]
        err = float((out - ref).abs().max())
        print(f"| {gm} | {qk} | {pv} | {passes[gm]} / {passes[qk]} / {passes[pv]} | {err:.2e} |", flush=True)
    # per-site relaxations on top of the shipped recipe (GEMMs x3, attention products one fp16 pass)
    print("\n| GEMM sites relaxed to 2 passes, recipe a16w16 (others x3; Q.K^T, P.V x1h) | max abs mel error vs fp64 |")
    print("|---|---|")
    r2 = "a16w16"
    for st in ({}, {"qkv": r2}, {"ffn1": r2}, {"ffn2": r2}, {"out": r2}, {"qkv": r2, "ffn1": r2},
               {"qkv": r2, "ffn1": r2, "ffn2": r2}, {"qkv": r2, "ffn1": r2, "ffn2": r2, "out": r2}):
        with Emu("x3", "x1h", "x1h", sites=st), torch.no_grad():
            out = O.forward(sd, hp, batch, inference=False, dtype=torch.float64)["mel"]
        print(f"| {' '.join(f'{k}={v}' for k, v in st.items()) or '(none)'} | {float((out - ref).abs().max()):.2e} |", flush=True)


if __name__ == "__main__":
    main()
