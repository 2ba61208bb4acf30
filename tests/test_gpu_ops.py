"""GPU parity of every C-ABI entry point against the oracle / fp64 torch on the same
seeded inputs.  Tolerances are stated per test; integer/byte work is bit-exact."""

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from lightningfastspeech2_b200 import ops
from oracle import fs2_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    g = np.random.default_rng(seed)
    return torch.from_numpy((g.standard_normal(shape) * scale).astype(np.float32))


@pytest.mark.parametrize("m,n,k,relu", [(300, 768, 256, False), (129, 80, 256, False), (1000, 1024, 256, True),
                                         (257, 256, 1024, False), (5, 2304, 768, False)])
def test_linear(m, n, k, relu):
    a, w, b = rnd(m, k, seed=1), rnd(n, k, seed=2, scale=k ** -0.5), rnd(n, seed=3, scale=0.1)
    ref = F.linear(a.double(), w.double(), b.double())
    ref = torch.relu(ref) if relu else ref
    got = ops.linear(a.to(DEV), w.to(DEV), b.to(DEV), relu=relu).cpu()
    assert (got - ref).abs().max() < 5e-5  # fp32 FMA accumulation over k<=1024


@pytest.mark.parametrize("bsz,t,d,n,ks", [(2, 37, 64, 128, 9), (3, 5, 32, 64, 3), (1, 130, 256, 1024, 9)])
def test_conv1d_dense(bsz, t, d, n, ks):
    x, w, b = rnd(bsz, t, d, seed=4), rnd(n, d, ks, seed=5, scale=(d * ks) ** -0.5), rnd(n, seed=6, scale=0.1)
    ref = torch.relu(F.conv1d(x.double().transpose(1, 2), w.double(), b.double(), padding=(ks - 1) // 2)).transpose(1, 2)
    wp = w.permute(0, 2, 1).reshape(n, ks * d).contiguous()
    got = ops.conv1d_dense(x.to(DEV), wp.to(DEV), b.to(DEV), ks, relu=True).cpu()
    assert (got - ref).abs().max() < 5e-5


@pytest.mark.parametrize("bsz,t,d,ks", [(2, 50, 256, 25), (3, 7, 64, 9), (1, 1, 32, 3), (2, 33, 768, 17)])
def test_dwconv1d(bsz, t, d, ks):
    x, w, b = rnd(bsz, t, d, seed=7), rnd(d, 1, ks, seed=8, scale=ks ** -0.5), rnd(d, seed=9, scale=0.1)
    ref = F.conv1d(x.double().transpose(1, 2), w.double(), b.double(), padding=(ks - 1) // 2, groups=d).transpose(1, 2)
    got = ops.dwconv1d(x.to(DEV), w[:, 0, :].t().contiguous().to(DEV), b.to(DEV)).cpu()
    assert (got - ref).abs().max() < 1e-5


@pytest.mark.parametrize("bsz,t,d,h", [(2, 150, 256, 2), (3, 70, 768, 2), (1, 64, 256, 2), (2, 129, 128, 2)])
def test_attention(bsz, t, d, h):
    qkv = rnd(bsz, t, 3 * d, seed=10)
    kpm = torch.zeros(bsz, t, dtype=torch.bool)
    kpm[-1, t // 2:] = True
    dh = d // h
    q, k, v = qkv.double().split(d, -1)
    hd = lambda z: z.reshape(bsz, t, h, dh).permute(0, 2, 1, 3)
    s = (hd(q) * dh ** -0.5) @ hd(k).transpose(-1, -2)
    s = s.masked_fill(kpm[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ hd(v)).permute(0, 2, 1, 3).reshape(bsz, t, d)
    got = ops.attention(qkv.to(DEV), kpm.to(DEV), h).cpu()
    assert (got - ref).abs().max() < 2e-5


def test_attention_fully_masked_row_is_nan_like_reference():
    qkv = rnd(2, 40, 768, seed=11)
    kpm = torch.zeros(2, 40, dtype=torch.bool)
    kpm[1] = True
    got = ops.attention(qkv.to(DEV), kpm.to(DEV), 2).cpu()
    assert torch.isnan(got[1]).all() and torch.isfinite(got[0]).all()


@pytest.mark.parametrize("m,d,res", [(100, 256, True), (33, 768, True), (7, 256, False), (5, 32, True)])
def test_add_layernorm(m, d, res):
    x, y, g, b = rnd(m, d, seed=12), rnd(m, d, seed=13), 1 + rnd(d, seed=14, scale=0.1), rnd(d, seed=15, scale=0.1)
    z = x.double() + (y.double() if res else 0)
    ref = F.layer_norm(z, (d,), g.double(), b.double(), 1e-5)
    got = ops.add_layernorm(x.to(DEV), y.to(DEV) if res else None, g.to(DEV), b.to(DEV)).cpu()
    assert (got - ref).abs().max() < 1e-5


def test_rowdot_mask():
    z, w, b = rnd(3, 41, 256, seed=16), rnd(1, 256, seed=17, scale=0.06), torch.tensor([1.7])
    mask = torch.zeros(3, 41, dtype=torch.bool)
    mask[1, 30:] = True
    ref = F.linear(z.double(), w.double(), b.double()).squeeze(-1).masked_fill(mask, 0)
    got = ops.rowdot_mask(z.to(DEV), w.to(DEV), b.to(DEV), mask.to(DEV)).cpu()
    assert (got - ref).abs().max() < 1e-5
    assert (got[mask] == 0).all()


def test_speaker_proj_and_front_end():
    from lightningfastspeech2_b200 import synthetic

    d = 256
    dv, w, b = rnd(5, 256, seed=18), rnd(d, 256, seed=19, scale=1 / 16), rnd(d, seed=20, scale=0.1)
    spk = ops.speaker_proj(dv.to(DEV), w.to(DEV), b.to(DEV))
    ref = torch.relu(F.linear(dv.double(), w.double(), b.double()))
    assert (spk.cpu() - ref).abs().max() < 1e-5
    emb = rnd(80, d, seed=21)
    emb[0] = 0
    pe = synthetic.sinusoid_table(5000, d)
    phones = torch.randint(1, 80, (5, 33))
    phones[2, 20:] = 0
    x, mask = ops.embed_pe_spk(phones.to(DEV), emb.to(DEV), pe.to(DEV), spk)
    refx = (emb[phones] + pe[:, :33]) + spk.cpu()[:, None, :]
    assert torch.equal(mask.cpu(), phones.eq(0))
    assert torch.equal(x.cpu(), refx)  # pure fp32 adds in the reference's association: bit-exact
    y = ops.add_pe_spk_(x.clone(), pe.to(DEV), spk)
    assert torch.equal(y.cpu(), (refx + pe[:, :33]) + spk.cpu()[:, None, :])


def test_duration_round_guard():
    g = np.random.default_rng(22)
    p = torch.from_numpy(g.uniform(-1.0, 2.5, (6, 40)).astype(np.float32))
    # exact .5 cases for round-half-even: exp(p)-1 == 0.5, 1.5, 2.5
    p[0, :3] = torch.log(torch.tensor([1.5, 2.5, 3.5]))
    mask = torch.zeros(6, 40, dtype=torch.bool)
    mask[1, 25:] = True
    p[2] = -3.0  # all-zero durations -> guard sets valid to 1
    mask[2, 10:] = True
    p[3, :] = -3.0
    p[3, :21] = 0.7  # sum 21*1 > 40//2 -> no guard ; boundary case next
    p[4, :] = -3.0
    p[4, :20] = 0.7  # sum 20 <= 20 -> guard
    ref = O.round_durations(p, mask)
    got = ops.duration_round_guard(p.to(DEV), mask.to(DEV)).cpu()
    assert got.dtype == torch.int32
    flips = (got != ref).sum().item()
    assert flips == 0, f"{flips} duration flips"


def test_bucket_embed_add():
    d, nb = 256, 256
    bins = torch.linspace(-2.5, 3.5, nb - 1)
    emb = rnd(nb, d, seed=23)
    val = rnd(4, 50, seed=24, scale=1.5)
    val[0, :5] = torch.tensor([-100.0, 100.0, float("nan"), float(bins[7] - 0.3) / 1.7, 0.0])
    val[0, 5] = (bins[100] - 0.3) / 1.7  # lands on/near a boundary after the affine map
    x = rnd(4, 50, d, seed=25)
    std, mean = 1.7, 0.3
    ref_idx = torch.bucketize(val * std + mean, bins)
    ref = x + emb[ref_idx]
    xg = x.to(DEV).clone()
    acc = torch.empty_like(xg)
    idx = ops.bucket_embed_add_(xg, val.to(DEV), std, mean, bins.to(DEV), emb.to(DEV), acc=acc, acc_init=True,
                                want_idx=True)
    assert torch.equal(idx.cpu(), ref_idx)
    assert torch.equal(xg.cpu(), ref)
    assert torch.equal(acc.cpu(), emb[ref_idx])
    forced = torch.randint(0, nb, (4, 50))
    xg2 = x.to(DEV).clone()
    ops.bucket_embed_add_(xg2, None, std, mean, bins.to(DEV), emb.to(DEV), idx_forced=forced.to(DEV), acc=acc)
    assert torch.equal(xg2.cpu(), x + emb[forced])
    assert torch.equal(acc.cpu(), emb[ref_idx] + emb[forced])


# ---- long depthwise kernels (the LightSpeech blocks use 13 ... 25 taps) on bench-sized launches ----
@pytest.mark.parametrize("ks", [11, 13, 17, 21, 23, 25])
@pytest.mark.parametrize("bsz,t,d", [(9, 2203, 256), (3, 2211, 768), (5, 70, 256), (2, 64, 256)])
def test_dwconv1d_long_kernels_large_launches(bsz, t, d, ks):
    """K >= 11 on launches of bench size: fp32 and planes inputs, fp32 / planes / fp16 outputs, ragged T, several channel
    blocks (d = 768), row limits"""
    import torch.nn.functional as F

    g = torch.Generator().manual_seed(ks * 1000 + d)
    x = torch.randn(bsz, t, d, generator=g)
    w = torch.randn(d, 1, ks, generator=g) * 0.2
    b = torch.randn(d, generator=g) * 0.1
    ref = F.conv1d(x.double().transpose(1, 2), w.double(), b.double(), padding=(ks - 1) // 2, groups=d).transpose(1, 2)
    wt = w[:, 0, :].t().contiguous().to(DEV)
    xd = x.to(DEV)
    got = ops.dwconv1d_planes(xd, wt, b.to(DEV), out="f32")
    assert float((got.cpu().double() - ref).abs().max()) < 2e-5
    xp = ops.split_bf16(xd)
    refp = F.conv1d(xp.float().cpu().double().transpose(1, 2), w.double(), b.double(), padding=(ks - 1) // 2,
                    groups=d).transpose(1, 2)
    gp = ops.dwconv1d_planes(xp, wt, b.to(DEV))
    assert float((gp.float().cpu().double() - refp).abs().max()) < 1e-4            # planes out: 2^-16 relative
    g32 = ops.dwconv1d_planes(xp, wt, b.to(DEV), out="f32")
    assert float((g32.cpu().double() - refp).abs().max()) < 2e-5
    gh = ops.dwconv1d_planes(xp, wt, b.to(DEV), out="f16")
    assert gh.lo is None and torch.equal(gh.hi, g32.half())
    # plane-form input of the 11 .. 23-tap kernels goes through the TMA-staged kernel (dwconv_tma.cu), fp32 input through
    # the per-thread-load kernel: same operations in the same order => the same bits
    same = ops.dwconv1d_planes(ops.merge_planes(xp), wt, b.to(DEV), out="f32")
    assert torch.equal(same, g32)
    if d == 256:   # row limits: kept rows are the unlimited run's bit for bit, input rows past the kept tiles read as zeros
        lens = torch.randint(1, t + 1, (bsz,), generator=g).to(torch.int32).to(DEV)
        extra = 20
        part = ops.dwconv1d_planes(xp, wt, b.to(DEV), row_limit=(lens, extra))
        for i, ln_ in enumerate(lens.tolist()):
            keep = min(t, (ln_ + extra + 127) // 128 * 128)
            inner = max(0, keep - (ks - 1) // 2)      # rows whose window stays inside the kept tiles
            assert torch.equal(part.hi[i, :inner], gp.hi[i, :inner]) and torch.equal(part.lo[i, :inner], gp.lo[i, :inner]), i
