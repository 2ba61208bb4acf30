"""tcgen05 attention (lfs2_attention_tc) against an fp64 torch restatement of
nn.MultiheadAttention's core (reference model.py:111-114 -> torch _sa_block): q scaled by
head_dim^-1/2, -inf on PAD keys, softmax over keys, P.V.  npass=3 is the fp32-parity mode
(tolerance 1e-4 abs on O(1) outputs), npass=1 the bf16 mode (3e-2)."""
import math

import numpy as np
import pytest
import torch

from lightningfastspeech2_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def ref_attention(qkv, kpm, nhead):
    b, t, d3 = qkv.shape
    d = d3 // 3
    dh = d // nhead
    q, k, v = qkv.double().split(d, dim=-1)
    q = q.view(b, t, nhead, dh).transpose(1, 2) / math.sqrt(dh)
    k = k.view(b, t, nhead, dh).transpose(1, 2)
    v = v.view(b, t, nhead, dh).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    if kpm is not None:
        s = s.masked_fill(kpm[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    return (p @ v).transpose(1, 2).reshape(b, t, d)


def make(b, t, d, seed, lens=None, scale=1.0):
    g = np.random.default_rng(seed)
    qkv = torch.from_numpy((g.standard_normal((b, t, 3 * d)) * scale).astype(np.float32))
    kpm = None
    if lens is not None:
        kpm = torch.arange(t)[None, :] >= torch.tensor(lens)[:, None]
    return qkv, kpm


@pytest.mark.parametrize("b,t,lens", [(2, 200, [200, 77]), (1, 64, None), (3, 1, None), (2, 130, [1, 129]),
                                      (2, 700, [700, 333]), (1, 128, [64]), (4, 257, [257, 256, 65, 3])])
@pytest.mark.parametrize("npass", [3, 1])
def test_matches_fp64(b, t, lens, npass):
    d, nhead = 256, 2
    qkv, kpm = make(b, t, d, seed=t + b, lens=lens)
    ref = ref_attention(qkv, kpm, nhead)
    planes = ops.split_bf16(qkv.to(DEV))
    ctx, cp = ops.attention_tc(planes, None if kpm is None else kpm.to(DEV), nhead, npass=npass, want_f32=True)
    tol = 1e-4 if npass == 3 else 3e-2
    err = (ctx.cpu() - ref).abs().max()
    assert err < tol, float(err)
    assert (cp.float().cpu() - ctx.cpu()).abs().max() < 1e-4


def test_large_logits_and_growing_max():
    """logit scale 6 => row maxima keep growing across key tiles: exercises the O rescale path"""
    qkv, kpm = make(2, 500, 256, seed=5, lens=[500, 410], scale=2.5)
    ref = ref_attention(qkv, kpm, 2)
    ctx, _ = ops.attention_tc(ops.split_bf16(qkv.to(DEV)), kpm.to(DEV), 2, want_f32=True)
    # |v| up to ~10 and logits of std 6: the split-bf16 products carry ~2^-16 relative error
    assert (ctx.cpu() - ref).abs().max() < 1e-3


def test_non_suffix_mask_and_fully_masked_utterance():
    qkv, _ = make(3, 150, 256, seed=6)
    g = torch.Generator().manual_seed(0)
    kpm = torch.rand(3, 150, generator=g) < 0.4
    kpm[1, :70] = True        # a whole leading key tile masked
    kpm[2, :] = True          # no valid key at all -> NaN like torch
    ref = ref_attention(qkv, kpm, 2)
    ctx, _ = ops.attention_tc(ops.split_bf16(qkv.to(DEV)), kpm.to(DEV), 2, want_f32=True)
    ctx = ctx.cpu()
    assert torch.isnan(ref[2]).all() and torch.isnan(ctx[2]).all()
    assert (ctx[:2] - ref[:2]).abs().max() < 1e-4


def test_agrees_with_fp32_simt_attention():
    qkv, kpm = make(2, 300, 256, seed=7, lens=[300, 190])
    a = ops.attention(qkv.to(DEV), kpm.to(DEV), 2)
    ctx, _ = ops.attention_tc(ops.split_bf16(qkv.to(DEV)), kpm.to(DEV), 2, want_f32=True)
    assert (ctx - a).abs().max() < 1e-4


def test_unsupported_head_dim_raises():
    qkv, _ = make(1, 32, 192, seed=8)
    with pytest.raises(NotImplementedError):
        ops.attention_tc(ops.split_bf16(qkv.to(DEV)), None, 2)


@pytest.mark.parametrize("npass", [3, 1])
def test_query_row_limit_skips_tiles_and_keeps_the_rest(npass):
    """lfs2_attention_tc_limited: query tiles starting at or after len + extra are not computed, all others are
    bit-identical to the unlimited launch"""
    b, t, d, nhead, extra = 3, 700, 256, 2, 28
    lens = [700, 100, 333]
    qkv, kpm = make(b, t, d, seed=11, lens=lens)
    planes = ops.split_bf16(qkv.to(DEV))
    full, _ = ops.attention_tc(planes, kpm.to(DEV), nhead, npass=npass, want_f32=True)
    lim = torch.tensor(lens, dtype=torch.int32, device=DEV)
    part, pp = ops.attention_tc(planes, kpm.to(DEV), nhead, npass=npass, want_f32=True, row_limit=(lim, extra))
    for i, n in enumerate(lens):
        keep = min(t, (n + extra + 127) // 128 * 128)
        assert torch.equal(part[i, :keep], full[i, :keep]), i
        assert torch.equal(pp.hi[i, :keep], part[i, :keep].bfloat16())


# ---- single-pass fp16 operands (compute mode "fp32" default: q, k, v as ONE fp16 plane, P as fp16)
@pytest.mark.parametrize("b,t,lens", [(2, 200, [200, 77]), (3, 1, None), (2, 700, [700, 333]), (4, 257, [257, 256, 65, 3])])
def test_fp16_operands_match_fp64(b, t, lens):
    d, nhead = 256, 2
    qkv, kpm = make(b, t, d, seed=t + b, lens=lens)
    q16 = qkv.half()
    ref = ref_attention(q16.float(), kpm, nhead)       # the fp16-rounded inputs are the kernel's inputs
    ctx, cp = ops.attention_tc(ops.Planes(q16.to(DEV), None), None if kpm is None else kpm.to(DEV), nhead, want_f32=True)
    err = (ctx.cpu() - ref).abs().max()
    assert err < 2e-3, float(err)                     # P in fp16: 2^-12 relative per probability, O(1) values
    assert (cp.float().cpu() - ctx.cpu()).abs().max() < 1e-4
    # and against the UNROUNDED inputs the error stays an order of magnitude below the bf16 mode's
    ref_full = ref_attention(qkv, kpm, nhead)
    bf = ops.attention_tc(ops.split_bf16(qkv.to(DEV)), None if kpm is None else kpm.to(DEV), nhead, npass=1, want_f32=True)[0]
    e16, ebf = (ctx.cpu() - ref_full).abs().max(), (bf.cpu() - ref_full).abs().max()
    print(f"attention t={t}: max err fp16 operands {float(e16):.2e}, bf16 operands {float(ebf):.2e}")
    assert e16 < 4e-3 and (t < 8 or e16 < 0.5 * ebf)


def test_fp16_operands_row_limit_and_masks():
    b, t, d, nhead = 3, 400, 256, 2
    lens = [400, 100, 333]
    qkv, kpm = make(b, t, d, seed=12, lens=lens)
    planes = ops.Planes(qkv.half().to(DEV), None)
    full, _ = ops.attention_tc(planes, kpm.to(DEV), nhead, want_f32=True)
    lim = torch.tensor(lens, dtype=torch.int32, device=DEV)
    part, _ = ops.attention_tc(planes, kpm.to(DEV), nhead, want_f32=True, row_limit=(lim, 28))
    for i, n in enumerate(lens):
        keep = min(t, (n + 28 + 127) // 128 * 128)
        assert torch.equal(part[i, :keep], full[i, :keep]), i


# ---- wide heads (head_dim 256 / 384: lfs2_attention_tc_wide; the 76 M configuration is d = 768, 2 heads)
@pytest.mark.parametrize("fmt,tol", [(torch.float16, 2e-3), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("b,t,d,lens", [(2, 200, 768, [200, 77]), (3, 1, 768, None), (2, 700, 768, [700, 333]),
                                        (4, 257, 768, [257, 256, 65, 3]), (2, 130, 512, [130, 64]), (1, 64, 768, None)])
def test_wide_heads_match_fp64(b, t, d, lens, fmt, tol):
    nhead = 2
    qkv, kpm = make(b, t, d, seed=t + b + d, lens=lens)
    q16 = qkv.to(fmt)
    ref = ref_attention(q16.float(), kpm, nhead)      # the rounded plane is the kernel's input
    ctx, cp = ops.attention_tc_wide(q16.to(DEV), None if kpm is None else kpm.to(DEV), nhead, want_f32=True)
    err = (ctx.cpu() - ref).abs().max()
    print(f"wide attention d={d} t={t} {fmt}: max err {float(err):.2e}")
    assert err < tol, float(err)
    assert (cp.float().cpu() - ctx.cpu()).abs().max() < 1e-4


def test_wide_heads_growing_max_masks_and_row_limit():
    b, t, d, nhead = 3, 500, 768, 2
    qkv, _ = make(b, t, d, seed=21, scale=1.6)    # logit std ~ 2.6 * sqrt(384)/sqrt(384): maxima keep growing -> O rescale path
    g = torch.Generator().manual_seed(1)
    kpm = torch.rand(b, t, generator=g) < 0.3
    kpm[1, :70] = True
    kpm[2, :] = True                              # no valid key at all -> NaN like torch
    q16 = qkv.half()
    ref = ref_attention(q16.float(), kpm, nhead)
    ctx, _ = ops.attention_tc_wide(q16.to(DEV), kpm.to(DEV), nhead, want_f32=True)
    ctx = ctx.cpu()
    assert torch.isnan(ref[2]).all() and torch.isnan(ctx[2]).all()
    assert (ctx[:2] - ref[:2]).abs().max() < 5e-3
    lens = [500, 100, 333]
    kpm2 = torch.arange(t)[None, :] >= torch.tensor(lens)[:, None]
    full, _ = ops.attention_tc_wide(q16.to(DEV), kpm2.to(DEV), nhead, want_f32=True)
    lim = torch.tensor(lens, dtype=torch.int32, device=DEV)
    part, _ = ops.attention_tc_wide(q16.to(DEV), kpm2.to(DEV), nhead, want_f32=True, row_limit=(lim, 28))
    for i, n in enumerate(lens):
        keep = min(t, (n + 28 + 127) // 128 * 128)
        assert torch.equal(part[i, :keep], full[i, :keep]), i


def test_wide_heads_agree_with_the_gemm_decomposed_attention():
    qkv, kpm = make(2, 300, 768, seed=22, lens=[300, 190])
    planes = ops.split_bf16(qkv.to(DEV))
    ref, _, _ = ops.attention_mat_fwd(planes, kpm.to(DEV), 2, npass=3)
    ctx, _ = ops.attention_tc_wide(qkv.half().to(DEV), kpm.to(DEV), 2, want_f32=True)
    assert (ctx - ref).abs().max() < 2e-3
