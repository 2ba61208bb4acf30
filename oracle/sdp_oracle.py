"""TEST INFRASTRUCTURE -- CPU restatement of the stochastic duration predictor's INFERENCE path.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this package; the product path never does.

Follows the reference at (paths relative to /root/reference/litfass):
  fastspeech2/model.py:196-207, 258-268, 299-309, 463-480   StochasticDurationPredictorWrapper and its call sites
  third_party/stochastic_duration_predictor/sdp.py:11-70     DilatedDepthSeparableConv
  .../sdp.py:73-95                                            ElementwiseAffine (reverse)
  .../sdp.py:98-164                                           ConvFlow (reverse)
  .../sdp.py:254-269, 330-349                                 StochasticDurationPredictor.forward(reverse=True)
  .../transforms.py:50-100, 103-212                           unconstrained rational-quadratic spline, inverse, linear tails
  .../normalization.py:5-28                                   LayerNorm over the channel dimension

Pinned by tests/golden/sdp_small.pt (outputs of the UNMODIFIED reference module with recorded noise,
oracle/make_goldens_sdp.py; replayed by tests/test_oracle_sdp.py).  Activations are kept channels-last (B, T, C) like
the CUDA path; the arithmetic per element is the reference's.  The training direction (the flows' negative log
likelihood, sdp.py:271-328) is not restated: the CUDA path does not implement it either."""
import math

import torch
import torch.nn.functional as F

NUM_BINS = 10
TAIL_BOUND = 5.0
MIN_W = MIN_H = MIN_D = 1e-3


def _ln(x, sd, key):
    return F.layer_norm(x, (x.shape[-1],), sd[key + ".gamma"].to(x.dtype), sd[key + ".beta"].to(x.dtype), 1e-5)


def dds_conv(x, valid, sd, pre, nlayers, g=None):
    """sdp.py:55-70 on channels-last x (B, T, C); valid (B, T, 1) = 1.0 on real phones"""
    if g is not None:
        x = x + g
    c = x.shape[-1]
    for i in range(nlayers):
        w = sd[f"{pre}.convs_sep.{i}.weight"].to(x.dtype)
        k = w.shape[-1]
        dil = k ** i
        y = F.conv1d((x * valid).transpose(1, 2), w, sd[f"{pre}.convs_sep.{i}.bias"].to(x.dtype),
                     padding=(k * dil - dil) // 2, dilation=dil, groups=c).transpose(1, 2)
        y = F.gelu(_ln(y, sd, f"{pre}.norms_1.{i}"))
        y = F.linear(y, sd[f"{pre}.convs_1x1.{i}.weight"][:, :, 0].to(x.dtype), sd[f"{pre}.convs_1x1.{i}.bias"].to(x.dtype))
        y = F.gelu(_ln(y, sd, f"{pre}.norms_2.{i}"))
        x = x + y
    return x * valid


def spline_inverse(y, uw, uh, ud):
    """inverse of the monotone rational-quadratic spline on [-5, 5] with identity tails (transforms.py:50-100 with
    inverse=True): y (...), uw / uh (..., 10), ud (..., 9) -> x (...)"""
    k = uw.shape[-1]
    inside = (y >= -TAIL_BOUND) & (y <= TAIL_BOUND)

    def knots(u, min_size):
        s = F.softmax(u, dim=-1)
        s = min_size + (1 - min_size * k) * s
        cum = F.pad(torch.cumsum(s, dim=-1), (1, 0))
        cum = 2 * TAIL_BOUND * cum - TAIL_BOUND
        cum[..., 0] = -TAIL_BOUND
        cum[..., -1] = TAIL_BOUND
        return cum, cum[..., 1:] - cum[..., :-1]

    cw, w = knots(uw, MIN_W)
    ch, h = knots(uh, MIN_H)
    edge = math.log(math.exp(1 - MIN_D) - 1)          # boundary derivative min_d + softplus(edge) = 1: C1 with the tails
    d = MIN_D + F.softplus(F.pad(ud, (1, 1), value=edge))
    search = ch.clone()
    search[..., -1] += 1e-6
    yc = torch.where(inside, y, torch.zeros_like(y))  # (values outside the interval pass through unchanged below)
    b = (torch.sum(yc[..., None] >= search, dim=-1) - 1).clamp(0, k - 1)[..., None]
    take = lambda t: t.gather(-1, b)[..., 0]
    cw_b, w_b, ch_b, h_b = take(cw), take(w), take(ch), take(h)
    delta = take(h / w)
    d0, d1 = take(d[..., :-1]), take(d[..., 1:])
    t = yc - ch_b
    s2 = d0 + d1 - 2 * delta
    qa = t * s2 + h_b * (delta - d0)
    qb = h_b * d0 - t * s2
    qc = -delta * t
    root = (2 * qc) / (-qb - torch.sqrt(qb * qb - 4 * qa * qc))
    return torch.where(inside, root * w_b + cw_b, y)


def conv_flow_reverse(z, valid, sd, pre, g, hidden):
    """sdp.py:140-164 with reverse=True on z (B, T, 2)"""
    x0, x1 = z[..., 0:1], z[..., 1]
    h = F.linear(x0, sd[pre + ".pre.weight"][:, :, 0].to(z.dtype), sd[pre + ".pre.bias"].to(z.dtype))
    h = dds_conv(h, valid, sd, pre + ".convs", 3, g=g)
    h = F.linear(h, sd[pre + ".proj.weight"][:, :, 0].to(z.dtype), sd[pre + ".proj.bias"].to(z.dtype)) * valid
    scale = 1.0 / math.sqrt(hidden)
    x1 = spline_inverse(x1, h[..., :NUM_BINS] * scale, h[..., NUM_BINS:2 * NUM_BINS] * scale, h[..., 2 * NUM_BINS:])
    return torch.stack([x0[..., 0], x1], dim=-1) * valid


def sdp_inference(x, src_mask, sd, pre, noise, noise_scale=1.0, dtype=torch.float32):
    """StochasticDurationPredictorWrapper(x, mask, inference=True) (model.py:476-480): x (B, T, d) encoder output, src_mask
    (B, T) True = PAD, noise (B, 2, T) the torch.randn draw of sdp.py:331 -> log-durations (B, T), 0 at PAD"""
    sd = {k[len(pre) + 1:]: v for k, v in sd.items() if k.startswith(pre + ".")}
    valid = (~src_mask).to(dtype)[..., None]
    x = x.to(dtype)
    hidden = sd["sdp.pre.weight"].shape[0]
    c = F.linear(x, sd["sdp.pre.weight"][:, :, 0].to(dtype), sd["sdp.pre.bias"].to(dtype))
    c = dds_conv(c, valid, sd, "sdp.convs", 3)
    c = F.linear(c, sd["sdp.proj.weight"][:, :, 0].to(dtype), sd["sdp.proj.bias"].to(dtype)) * valid
    nflows = len({k.split(".")[2] for k in sd if k.startswith("sdp.flows.")})
    order = list(range(nflows))[::-1]
    order = order[:-2] + [order[-1]]                   # sdp.py:331: the flow next to the affine layer is dropped
    z = (noise.to(dtype) * noise_scale).transpose(1, 2)  # (B, T, 2)
    for j in order:
        z = torch.flip(z, [-1])
        if j == 0:                                       # ElementwiseAffine, reverse (sdp.py:93-95)
            tr = sd["sdp.flows.0.translation"][:, 0].to(dtype)
            ls = sd["sdp.flows.0.log_scale"][:, 0].to(dtype)
            z = (z - tr) * torch.exp(-ls) * valid
        else:
            z = conv_flow_reverse(z, valid, sd, f"sdp.flows.{j}", c, hidden)
    return z[..., 0].masked_fill(src_mask, 0.0)


def stochastic_durations(logw, src_mask):
    """model.py:302-309 for duration_stochastic=True: ceil(exp(logw + 1e-9)), 0 where logw == 0, clamp, int, and the
    all-ones guard"""
    dur = torch.ceil(torch.exp(logw + 1e-9)).masked_fill(logw == 0, 0)
    dur = torch.clamp(dur, min=0).int()
    for i in range(len(dur)):
        if dur[i][~src_mask[i]].sum() <= (~src_mask[i]).sum() // 2:
            dur[i][~src_mask[i]] = 1
    return dur
