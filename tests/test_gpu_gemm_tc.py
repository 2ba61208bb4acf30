"""tcgen05 GEMM / Conv1d (lfs2_gemm_tc) against fp64 torch on the same seeded inputs.

npass=3 (bf16 hi/lo split, three MMAs per k-step) is the fp32-parity mode: tolerance 2e-4
abs on O(1) outputs with K <= 1024 (theory: ~2^-16 relative per product).  npass=1 is the
bf16 mode: tolerance 3e-2."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from lightningfastspeech2_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    g = np.random.default_rng(seed)
    return torch.from_numpy((g.standard_normal(shape) * scale).astype(np.float32))


def test_split_bf16_reconstructs_to_2e_minus_16():
    x = rnd(1000, 256, seed=1, scale=3.0).to(DEV)
    p = ops.split_bf16(x)
    rel = ((p.float() - x).abs() / x.abs().clamp_min(1e-20)).max()
    assert rel < 2.0 ** -15
    assert torch.equal(p.hi, x.to(torch.bfloat16))


@pytest.mark.parametrize("m,n,k", [(128, 256, 256), (300, 768, 256), (1000, 1024, 256), (257, 256, 1024),
                                   (129, 80, 256), (5, 2304, 768), (20000, 256, 256)])
@pytest.mark.parametrize("npass", [3, 1])
def test_linear(m, n, k, npass):
    a, w, b = rnd(m, k, seed=1), rnd(n, k, seed=2, scale=k ** -0.5), rnd(n, seed=3, scale=0.1)
    ref = F.linear(a.double(), w.double(), b.double())
    ap, wp = ops.split_bf16(a.to(DEV)), ops.split_bf16(w.to(DEV))
    out = ops.gemm_tc(ap, wp, b.to(DEV), npass=npass)                       # fp32 out (TMA store, 128B swizzle)
    planes = ops.gemm_tc(ap, wp, b.to(DEV), npass=npass, out="planes")      # hi/lo planes out (64B swizzle)
    tol = 2e-4 if npass == 3 else 3e-2
    err = (out.cpu() - ref).abs().max()
    assert err < tol, float(err)
    assert (planes.float().cpu() - out.cpu()).abs().max() < 1e-4
    assert torch.equal(planes.hi, out.to(torch.bfloat16))


def test_relu_and_planes_only():
    a, w, b = rnd(333, 256, seed=4), rnd(1024, 256, seed=5, scale=1 / 16), rnd(1024, seed=6, scale=0.1)
    ref = torch.relu(F.linear(a.double(), w.double(), b.double()))
    planes = ops.gemm_tc(ops.split_bf16(a.to(DEV)), ops.split_bf16(w.to(DEV)), b.to(DEV), relu=True, out="planes")
    assert (planes.float().cpu() - ref).abs().max() < 2e-4


@pytest.mark.parametrize("bsz,t,d,n,ks", [(2, 37, 64, 128, 9), (3, 130, 256, 1024, 9), (1, 5, 32, 256, 3),
                                          (4, 260, 256, 256, 3)])
def test_conv1d_taps(bsz, t, d, n, ks):
    x, w, b = rnd(bsz, t, d, seed=7), rnd(n, d, ks, seed=8, scale=(d * ks) ** -0.5), rnd(n, seed=9, scale=0.1)
    ref = F.conv1d(x.double().transpose(1, 2), w.double(), b.double(), padding=(ks - 1) // 2).transpose(1, 2)
    wp = w.permute(0, 2, 1).reshape(n, ks * d).contiguous()
    out = ops.gemm_tc(ops.split_bf16(x.to(DEV)), ops.split_bf16(wp.to(DEV)), b.to(DEV), taps=ks)
    err = (out.cpu() - ref).abs().max()
    assert err < 2e-4, float(err)


@pytest.mark.parametrize("m,k,relu,res", [(300, 256, False, True), (1000, 1024, False, True), (260, 256, True, False)])
def test_layernorm_epilogue(m, k, relu, res):
    n = 256
    a, w, b = rnd(m, k, seed=10), rnd(n, k, seed=11, scale=k ** -0.5), rnd(n, seed=12, scale=0.1)
    r = rnd(m, n, seed=13)
    g, bt = 1 + rnd(n, seed=14, scale=0.1), rnd(n, seed=15, scale=0.1)
    v = F.linear(a.double(), w.double(), b.double())
    if relu:
        v = torch.relu(v)
    if res:
        v = v + r.double()
    ref = F.layer_norm(v, (n,), g.double(), bt.double(), 1e-5)
    kw = dict(relu=relu, residual=ops.split_bf16(r.to(DEV)) if res else None, gamma=g.to(DEV), beta=bt.to(DEV))
    ap, wp = ops.split_bf16(a.to(DEV)), ops.split_bf16(w.to(DEV))
    out = ops.gemm_tc(ap, wp, b.to(DEV), **kw)
    planes = ops.gemm_tc(ap, wp, b.to(DEV), out="planes", **kw)
    err = (out.cpu() - ref).abs().max()
    assert err < 3e-4, float(err)
    assert (planes.float().cpu() - out.cpu()).abs().max() < 1e-4


def test_matches_fp32_simt_gemm_closely():
    """same inputs through the exact-fp32 CUDA-core GEMM: the split path must agree to ~1e-5"""
    a, w, b = rnd(4096, 256, seed=16), rnd(768, 256, seed=17, scale=1 / 16), rnd(768, seed=18, scale=0.1)
    ref = ops.linear(a.to(DEV), w.to(DEV), b.to(DEV))
    out = ops.gemm_tc(ops.split_bf16(a.to(DEV)), ops.split_bf16(w.to(DEV)), b.to(DEV))
    assert (out - ref).abs().max() < 1e-4


def test_bf16_mode_keeps_the_residual_in_fp32_precision():
    """npass=1 rounds the GEMM operands to bf16 but the residual still rides as hi+lo planes"""
    m, k, n = 500, 256, 256
    a, w, b = rnd(m, k, seed=20, scale=1e-3), rnd(n, k, seed=21, scale=k ** -0.5), rnd(n, seed=22, scale=1e-3)
    r = rnd(m, n, seed=23)
    g, bt = torch.ones(n), torch.zeros(n)
    ref = F.layer_norm(F.linear(a.double(), w.double(), b.double()) + r.double(), (n,), g.double(), bt.double(), 1e-5)
    out = ops.gemm_tc(ops.split_bf16(a.to(DEV)), ops.split_bf16(w.to(DEV)), b.to(DEV), residual=ops.split_bf16(r.to(DEV)),
                      gamma=g.to(DEV), beta=bt.to(DEV), npass=1)
    assert (out.cpu() - ref).abs().max() < 1e-4


def test_merge_and_dwconv_on_planes():
    x = rnd(3, 70, 256, seed=24).to(DEV)
    xp = ops.split_bf16(x)
    assert (ops.merge_planes(xp) - x).abs().max() < 2.0 ** -15 * x.abs().max()
    wt, bias = rnd(9, 256, seed=25, scale=0.3).to(DEV), rnd(256, seed=26, scale=0.1).to(DEV)
    ref = ops.dwconv1d(x, wt, bias)
    got = ops.dwconv1d_planes(xp, wt, bias)              # planes in -> planes out
    got32 = ops.dwconv1d_planes(x, wt, bias, out="f32")  # fp32 in -> fp32 out
    assert torch.equal(got32, ref)
    assert (got.float() - ref).abs().max() < 1e-4


@pytest.mark.parametrize("npass,tol", [(3, 5e-5), (1, 3e-2)])
@pytest.mark.parametrize("m,f", [(128, 256), (1000, 1024), (40000, 1024), (333, 2048)])
def test_ffn_fused_tc(m, f, npass, tol):
    """LayerNorm(res + relu(u.W1^T + b1).W2^T + b2) in one kernel vs an fp64 restatement"""
    import torch.nn.functional as F

    from lightningfastspeech2_b200 import ops

    g = torch.Generator().manual_seed(m + f)
    d = 256
    u = torch.randn(m, d, generator=g)
    res = torch.randn(m, d, generator=g)
    w1 = torch.randn(f, d, generator=g) / d ** 0.5
    w2 = torch.randn(d, f, generator=g) / f ** 0.5
    b1, b2 = torch.randn(f, generator=g) * 0.1, torch.randn(d, generator=g) * 0.1
    gam, bet = 1 + 0.1 * torch.randn(d, generator=g), 0.1 * torch.randn(d, generator=g)
    ref = F.layer_norm(res.double() + torch.relu(u.double() @ w1.double().t() + b1.double()) @ w2.double().t() + b2.double(),
                       (d,), gam.double(), bet.double(), 1e-5)
    dev = "cuda"
    sp = lambda t: ops.split_bf16(t.to(dev).contiguous())
    out = ops.ffn_fused_tc(sp(u), sp(w1), b1.to(dev), sp(w2), b2.to(dev), sp(res), gam.to(dev), bet.to(dev), 1e-5, npass=npass)
    err = (out.float().cpu().double() - ref).abs().max().item()
    print(f"ffn_fused m={m} f={f} npass={npass}: max err {err:.3e}")
    assert err < tol, err


@pytest.mark.parametrize("bsz,t,lens,extra", [(3, 300, [300, 1, 130], 0), (4, 257, [0, 128, 100, 257], 28), (2, 90, [40, 90], 5)])
def test_ffn_fused_tc_row_limit_keeps_needed_tiles(bsz, t, lens, extra):
    """row-limited launch: every row t_row < roundup128(len + extra) of every utterance must equal the full launch bit
    for bit (tiles are 128 flattened rows, so a kept tile may also carry rows of the neighbours: those are unspecified)"""
    from lightningfastspeech2_b200 import ops

    g = torch.Generator().manual_seed(bsz * t)
    d, f, dev = 256, 1024, "cuda"
    sp = lambda x: ops.split_bf16(x.to(dev).contiguous())
    u, res = sp(torch.randn(bsz, t, d, generator=g)), sp(torch.randn(bsz, t, d, generator=g))
    w1, w2 = sp(torch.randn(f, d, generator=g) / 16), sp(torch.randn(d, f, generator=g) / 32)
    b1, b2 = torch.randn(f, generator=g).to(dev) * 0.1, torch.randn(d, generator=g).to(dev) * 0.1
    gam, bet = torch.ones(d, device=dev), torch.zeros(d, device=dev)
    full = ops.ffn_fused_tc(u, w1, b1, w2, b2, res, gam, bet)
    lim = torch.tensor(lens, dtype=torch.int32, device=dev)
    cache = {}
    part = ops.ffn_fused_tc(u, w1, b1, w2, b2, res, gam, bet, row_limit=(lim, extra, cache))
    again = ops.ffn_fused_tc(u, w1, b1, w2, b2, res, gam, bet, row_limit=(lim, extra, cache))  # reuses the tile list
    assert ("ffn", bsz, t) in cache
    for b, n in enumerate(lens):
        keep = min(t, (n + extra + 127) // 128 * 128)
        for o in (part, again):
            assert torch.equal(o.hi[b, :keep], full.hi[b, :keep]) and torch.equal(o.lo[b, :keep], full.lo[b, :keep]), b


@pytest.mark.parametrize("npass", [3, 1])
def test_two_cta_multicast_path(npass):
    """launches with >= 296 row tiles and 256-column tiles run as 2-CTA clusters that multicast the weight slabs:
    odd tile counts (the tail pairs with an out-of-range tile), several column tiles, LayerNorm + residual, row limits"""
    tol = 2e-4 if npass == 3 else 3e-2
    m = 305 * 128 - 77  # 305 row tiles: odd
    a, w, b = rnd(m, 256, seed=21), rnd(768, 256, seed=22, scale=1 / 16), rnd(768, seed=23, scale=0.1)
    ap, wp = ops.split_bf16(a.to(DEV)), ops.split_bf16(w.to(DEV))
    out = ops.gemm_tc(ap, wp, b.to(DEV), npass=npass)
    ref = F.linear(a.double(), w.double(), b.double())
    assert (out.cpu() - ref).abs().max() < tol
    # LayerNorm epilogue with the residual on the tensor core
    n = 256
    w2, b2, r = rnd(n, 256, seed=24, scale=1 / 16), rnd(n, seed=25, scale=0.1), rnd(m, n, seed=26)
    g, bt = 1 + rnd(n, seed=27, scale=0.1), rnd(n, seed=28, scale=0.1)
    ref = F.layer_norm(F.linear(a.double(), w2.double(), b2.double()) + r.double(), (n,), g.double(), bt.double(), 1e-5)
    out = ops.gemm_tc(ap, ops.split_bf16(w2.to(DEV)), b2.to(DEV), residual=ops.split_bf16(r.to(DEV)), gamma=g.to(DEV),
                      beta=bt.to(DEV), npass=npass, out="planes")
    assert (out.float().cpu() - ref).abs().max() < (2e-4 if npass == 3 else 3e-2)
    # row-limited launch over (B, T, d): kept rows equal the unlimited launch bit for bit
    bsz, t = 64, 700
    x = ops.split_bf16(rnd(bsz, t, 256, seed=29).to(DEV))
    lens = torch.randint(1, t + 1, (bsz,), generator=torch.Generator().manual_seed(3)).to(torch.int32)
    full = ops.gemm_tc(x, wp, b.to(DEV), npass=npass, out="planes")
    part = ops.gemm_tc(x, wp, b.to(DEV), npass=npass, out="planes", row_limit=(lens.to(DEV), 28, {}))
    for i, ln_ in enumerate(lens.tolist()):
        keep = min(t, (ln_ + 28 + 127) // 128 * 128)
        assert torch.equal(part.hi[i, :keep], full.hi[i, :keep]) and torch.equal(part.lo[i, :keep], full.lo[i, :keep]), i


# ---- lfs2_gemm_tc_ex: fp16 output plane, dilation, leaky ReLU, residual without LayerNorm, row mask, 64-column tiles
@pytest.mark.parametrize("m,n,k", [(300, 768, 256), (129, 80, 256), (4000, 768, 256)])
def test_f16_output_plane(m, n, k):
    a, w, b = rnd(m, k, seed=31), rnd(n, k, seed=32, scale=k ** -0.5), rnd(n, seed=33, scale=0.1)
    ref = F.linear(a.double(), w.double(), b.double())
    p = ops.gemm_tc(ops.split_bf16(a.to(DEV)), ops.split_bf16(w.to(DEV)), b.to(DEV), out="f16")
    assert p.lo is None and p.hi.dtype == torch.float16 and tuple(p.hi.shape) == (m, n)
    # the fp32 accumulator rounded once to fp16: half an ulp of 11 significant bits
    assert ((p.hi.double().cpu() - ref).abs() <= 2.0 ** -11 * ref.abs() + 2e-4).all()
    f32 = ops.gemm_tc(ops.split_bf16(a.to(DEV)), ops.split_bf16(w.to(DEV)), b.to(DEV))
    assert torch.equal(p.hi, f32.half())


def test_f16_output_saturates_instead_of_overflowing():
    a, w = torch.full((128, 32), 300.0), torch.full((16, 32), 300.0)
    p = ops.gemm_tc(ops.split_bf16(a.to(DEV)), ops.split_bf16(w.to(DEV)), None, out="f16")
    assert torch.isfinite(p.hi).all() and float(p.hi.max()) == 65504.0


@pytest.mark.parametrize("bsz,t,c,n,ks,dil", [(2, 300, 64, 64, 3, 5), (1, 513, 128, 128, 7, 3), (3, 200, 32, 32, 11, 5),
                                              (2, 140, 256, 256, 3, 1), (1, 77, 512, 256, 7, 1)])
def test_dilated_conv_leaky_relu_and_residual(bsz, t, c, n, ks, dil):
    """HiFi-GAN ResBlock pieces (third_party/hifigan/models.py:86-93): dilated 'same' conv, leaky ReLU epilogue,
    x + conv(.) with the residual on the tensor core, narrow channel counts (64-column tiles)"""
    x, w, b = rnd(bsz, t, c, seed=41), rnd(n, c, ks, seed=42, scale=(c * ks) ** -0.5), rnd(n, seed=43, scale=0.1)
    conv = F.conv1d(x.double().transpose(1, 2), w.double(), b.double(), padding=dil * (ks - 1) // 2,
                    dilation=dil).transpose(1, 2)
    wp = ops.split_bf16(w.permute(0, 2, 1).reshape(n, ks * c).contiguous().to(DEV))
    xp = ops.split_bf16(x.to(DEV))
    out = ops.gemm_tc(xp, wp, b.to(DEV), taps=ks, dilation=dil, leaky_slope=0.1)
    assert (out.cpu() - F.leaky_relu(conv, 0.1)).abs().max() < 2e-4
    if n == c:
        res = ops.gemm_tc(xp, wp, b.to(DEV), taps=ks, dilation=dil, residual=xp, out="planes")
        assert (res.float().cpu() - (conv + x.double())).abs().max() < 3e-4


def test_row_mask_writes_zero_rows():
    bsz, t, c, n = 3, 260, 64, 128
    x, w, b = rnd(bsz, t, c, seed=44), rnd(n, c, seed=45, scale=c ** -0.5), rnd(n, seed=46)
    lens = torch.tensor([260, 100, 1])
    mask = (torch.arange(t)[None, :] >= lens[:, None]).to(DEV)
    full = ops.gemm_tc(ops.split_bf16(x.to(DEV)), ops.split_bf16(w.to(DEV)), b.to(DEV), taps=1, row_mask=None)
    out = ops.gemm_tc(ops.split_bf16(x.to(DEV)), ops.split_bf16(w.to(DEV)), b.to(DEV), taps=1, row_mask=mask)
    pl = ops.gemm_tc(ops.split_bf16(x.to(DEV)), ops.split_bf16(w.to(DEV)), b.to(DEV), taps=1, row_mask=mask, out="planes")
    full = full.view(bsz, t, n)
    assert torch.equal(out[~mask], full[~mask]) and float(out[mask].abs().max()) == 0.0
    assert float(pl.hi[mask].abs().max()) == 0.0 and float(pl.lo[mask].abs().max()) == 0.0


# ---- npass = 2: ONE fp16 activation plane against the bf16 hi/lo weight planes (a.w_hi + a.w_lo) -------------------
def _w_planes_value(w):
    """what the fp16 hi/lo weight planes of the 2-pass recipe hold (ops.split_f16), fp64"""
    hi = w.half()
    return hi.double() + (w - hi.float()).half().double()


def test_split_f16_planes():
    w = rnd(512, 256, seed=60, scale=0.05).to(DEV)
    p = ops.split_f16(w)
    assert p.hi.dtype == torch.float16 and torch.equal(p.hi, w.half())
    assert torch.equal(p.hi.double() + p.lo.double(), _w_planes_value(w.cpu()).to(DEV))
    assert ((p.hi.double() + p.lo.double() - w.double()).abs() <= 2.0 ** -22 * w.double().abs() + 2.0 ** -25).all()


@pytest.mark.parametrize("m,n,k,taps", [(300, 768, 256, 1), (129, 80, 256, 1), (20000, 768, 256, 1), (260, 256, 64, 3)])
@pytest.mark.parametrize("out", ["f32", "f16", "planes"])
def test_two_pass_fp16_activation_plane(m, n, k, taps, out):
    """the 2-pass product is exact in its operands: result == fp16(a) . (w_hi + w_lo) to the fp32 accumulator's rounding; against the un-rounded product the error is the fp16 rounding of a (2^-12 relative)"""
    a, w, b = rnd(2, m // 2, k, seed=61), rnd(n, taps * k, seed=62, scale=(taps * k) ** -0.5), rnd(n, seed=63, scale=0.1)
    a16 = a.half()
    wv = _w_planes_value(w)
    x = a16.double()
    if taps == 1:
        ref = x @ wv.t() + b.double()
    else:
        ref = F.conv1d(x.transpose(1, 2), wv.view(n, taps, k).permute(0, 2, 1).contiguous(), b.double(),
                       padding=(taps - 1) // 2).transpose(1, 2)
    ap = ops.split_bf16(a.to(DEV), want_f16=True)
    assert torch.equal(ap.h.cpu(), a16)
    got = ops.gemm_tc(ap, ops.split_f16(w.to(DEV)), b.to(DEV), taps=taps, npass=2, out=out)
    if out == "f32":
        assert (got.double().cpu() - ref).abs().max() < 3e-6 * max(1.0, float(ref.abs().max()))
    elif out == "f16":
        assert got.lo is None and got.hi.dtype == torch.float16
        assert ((got.hi.double().cpu() - ref).abs() <= 2.0 ** -11 * ref.abs() + 1e-5).all()
    else:
        assert (got.float().double().cpu() - ref).abs().max() < 2.0 ** -15 * max(1.0, float(ref.abs().max()))
    exact = F.linear(a.double(), w.double(), b.double()) if taps == 1 else None
    if exact is not None and out == "f32":
        assert (got.double().cpu() - exact).abs().max() < 1.5e-3   # 11-bit activations, K = 256: ~2^-12 * sqrt(K) * |a||w|


def test_two_pass_rejects_layernorm_and_missing_plane():
    a, w = rnd(128, 256, seed=64), rnd(256, 256, seed=65, scale=1 / 16)
    ap, wp = ops.split_bf16(a.to(DEV), want_f16=True), ops.split_f16(w.to(DEV))
    with pytest.raises(Exception):
        ops.gemm_tc(ap, wp, None, gamma=torch.ones(256, device=DEV), beta=torch.zeros(256, device=DEV), out="planes", npass=2)
    with pytest.raises(ValueError):
        ops.gemm_tc(ops.split_bf16(a.to(DEV)), wp, None, npass=2)


@pytest.mark.parametrize("m,f", [(128, 256), (1000, 1024), (40000, 1024), (333, 2048)])
def test_ffn_fused_tc_two_pass(m, f):
    """npass = 2: u as one fp16 plane, the intermediate packed as fp16, weights as hi/lo planes -- against an fp64
    restatement that rounds exactly those operands (tight), and against the un-rounded fp64 result (the recipe's cost);
    the fp16 output plane is the rounded hi + lo result"""
    g = torch.Generator().manual_seed(m + f + 1)
    d = 256
    u = torch.randn(m, d, generator=g)
    res = torch.randn(m, d, generator=g)
    w1 = torch.randn(f, d, generator=g) / d ** 0.5
    w2 = torch.randn(d, f, generator=g) / f ** 0.5
    b1, b2 = torch.randn(f, generator=g) * 0.1, torch.randn(d, generator=g) * 0.1
    gam, bet = 1 + 0.1 * torch.randn(d, generator=g), 0.1 * torch.randn(d, generator=g)
    ln = lambda z: F.layer_norm(z, (d,), gam.double(), bet.double(), 1e-5)
    exact = ln(res.double() + torch.relu(u.double() @ w1.double().t() + b1.double()) @ w2.double().t() + b2.double())
    v = torch.relu(u.half().double() @ _w_planes_value(w1).t() + b1.double())
    emu = ln(res.double() + v.float().half().double() @ _w_planes_value(w2).t() + b2.double())
    sp = lambda t: ops.split_bf16(t.to(DEV).contiguous())
    wt1, bias1 = torch.zeros(1, d, device=DEV), torch.zeros(d, device=DEV)
    wt1[0] = 1.0
    up = ops.dwconv1d_planes(u.to(DEV).view(1, m, d), wt1, bias1, out="f16")     # identity depthwise conv -> fp16 plane of u
    assert up.lo is None and torch.equal(up.hi.view(m, d).cpu(), u.half())
    up = ops.Planes(up.hi.view(m, d), None)
    sp16 = lambda t: ops.split_f16(t.to(DEV).contiguous())
    out = ops.ffn_fused_tc(up, sp16(w1), b1.to(DEV), sp16(w2), b2.to(DEV), sp(res), gam.to(DEV), bet.to(DEV), 1e-5, npass=2,
                           want_f16=True)
    got = out.float().cpu().double()
    e_emu, e_exact = (got - emu).abs().max().item(), (got - exact).abs().max().item()
    print(f"ffn_fused two-pass m={m} f={f}: vs operand-rounded fp64 {e_emu:.3e}, vs exact {e_exact:.3e}")
    # v sits at an fp16 rounding boundary in a few places (fp32 vs fp64 accumulation): a flipped v moves one output by up
    # to ulp(v) * |w2| ~ 3e-4, so the max is loose -- but flips are sparse, the rms stays at fp32-accumulation level
    # (a dropped lo weight plane would show as ~1e-4 rms)
    rms = float((got - emu).pow(2).mean().sqrt())
    assert e_emu < 1e-3 and rms < 1.5e-5 and e_exact < 3e-3, (e_emu, rms, e_exact)
    assert ((out.h.float() - out.float()).abs() <= 2.0 ** -11 * out.float().abs() + 1e-7).all()   # fp16 of the fp32 row
    # the 3-pass kernel with the same residual path (32 x 32 identity block) stays at fp32 parity
    o3 = ops.ffn_fused_tc(sp(u), sp(w1), b1.to(DEV), sp(w2), b2.to(DEV), sp(res), gam.to(DEV), bet.to(DEV), 1e-5, npass=3)
    assert (o3.float().cpu().double() - exact).abs().max() < 1e-4
