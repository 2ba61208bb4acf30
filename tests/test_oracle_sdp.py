"""oracle/sdp_oracle.py (stochastic duration predictor, inference direction) against the golden the UNMODIFIED reference
produced (oracle/make_goldens_sdp.py): same weights (rebuilt from the seed), same inputs, the recorded noise draw."""
import os

import torch

from lightningfastspeech2_b200 import synthetic
from oracle import sdp_oracle as S

GOLD = os.path.join(os.path.dirname(__file__), "golden", "sdp_small.pt")


def test_oracle_matches_the_reference_module():
    g = torch.load(GOLD, weights_only=False)
    sd = {"dp." + k: v for k, v in synthetic.fill_state_dict(g["shapes"], seed=g["seed"]).items()}
    for case in g["cases"]:
        logw = S.sdp_inference(case["x"], case["src_mask"], sd, "dp", case["noise"], noise_scale=case["sigma"])
        assert logw.shape == case["logw"].shape
        err = float((logw - case["logw"]).abs().max())
        assert err < 2e-5, err                       # fp32 on both sides: conv / softmax / cumsum ordering
        assert torch.equal(S.stochastic_durations(case["logw"], case["src_mask"]), case["duration_rounded"])
        l64 = S.sdp_inference(case["x"], case["src_mask"], sd, "dp", case["noise"], noise_scale=case["sigma"], dtype=torch.float64)
        assert float((l64 - case["logw"].double()).abs().max()) < 2e-5


def test_spline_inverse_is_the_inverse_of_the_forward_spline():
    """property check independent of the golden: x -> forward spline (closed form) -> inverse gives x back; values
    outside [-5, 5] pass through"""
    g = torch.Generator().manual_seed(3)
    n, k = 500, S.NUM_BINS
    uw, uh, ud = torch.randn(n, k, generator=g, dtype=torch.float64), torch.randn(n, k, generator=g, dtype=torch.float64), \
        torch.randn(n, k - 1, generator=g, dtype=torch.float64)
    x = (torch.rand(n, generator=g, dtype=torch.float64) * 2 - 1) * 4.9

    # forward direction written out from the knots (monotone rational-quadratic segment)
    import math
    import torch.nn.functional as F

    def knots(u, m):
        s = m + (1 - m * k) * F.softmax(u, -1)
        c = F.pad(torch.cumsum(s, -1), (1, 0)) * 10 - 5
        c[..., 0], c[..., -1] = -5.0, 5.0
        return c
    cw, ch = knots(uw, 1e-3), knots(uh, 1e-3)
    d = 1e-3 + F.softplus(F.pad(ud, (1, 1), value=math.log(math.exp(1 - 1e-3) - 1)))
    b = (torch.sum(x[:, None] >= cw, -1) - 1).clamp(0, k - 1)[:, None]
    t = lambda z: z.gather(-1, b)[:, 0]
    w_b, h_b = t(cw[:, 1:] - cw[:, :-1]), t(ch[:, 1:] - ch[:, :-1])
    delta, d0, d1 = h_b / w_b, t(d[:, :-1]), t(d[:, 1:])
    th = (x - t(cw[:, :-1])) / w_b
    y = t(ch[:, :-1]) + h_b * (delta * th * th + d0 * th * (1 - th)) / (delta + (d0 + d1 - 2 * delta) * th * (1 - th))
    back = S.spline_inverse(y, uw, uh, ud)
    assert float((back - x).abs().max()) < 1e-9
    far = torch.tensor([-7.5, 5.0001, 9.0], dtype=torch.float64)
    assert torch.equal(S.spline_inverse(far, uw[:3], uh[:3], ud[:3]), far)
