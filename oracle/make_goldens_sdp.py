#!/usr/bin/env python
"""Golden vectors of the stochastic duration predictor (inference direction) from the UNMODIFIED reference (authoring
container only):

    python oracle/make_goldens_sdp.py      # writes tests/golden/sdp_small.pt

The reference's StochasticDurationPredictorWrapper (litfass/fastspeech2/model.py:463-480 around
third_party/stochastic_duration_predictor/sdp.py, imported through oracle/ref_shim.py) gets seeded weights and runs on
CPU with inference=True; its one random draw (torch.randn, sdp.py:331) is RECORDED.  The golden stores inputs, the draw,
the predicted log-durations and the durations model.py:302-309 derives from them; the weights are regenerated from the
seed wherever it is replayed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lightningfastspeech2_b200 import synthetic  # noqa: E402
from oracle import ref_shim  # noqa: E402

CFG = dict(nlayers=3, in_channels=256, filter_size=256, kernel_size=3, dropout=0.0)


def main():
    if hasattr(ref_shim, "install"):
        ref_shim.install()
    import litfass.fastspeech2.model as rm

    mod = rm.StochasticDurationPredictorWrapper(CFG["nlayers"], CFG["in_channels"], CFG["filter_size"], CFG["kernel_size"],
                                                CFG["dropout"])
    seed = 9
    sd = synthetic.fill_state_dict(mod.state_dict(), seed=seed)
    mod.load_state_dict(sd)
    mod.eval()
    g = torch.Generator().manual_seed(31)
    cases = []
    for bsz, tp, lens, sigma in ((2, 13, [13, 8], 1.0), (3, 40, [40, 1, 27], 0.667)):
        x = torch.randn(bsz, tp, CFG["in_channels"], generator=g)
        src_mask = torch.arange(tp)[None, :] >= torch.tensor(lens)[:, None]
        draws = []
        orig = torch.randn

        def rec_randn(*size, **kw):
            z = orig(*size, generator=g)
            draws.append(z.clone())
            return z

        torch.randn = rec_randn
        try:
            with torch.no_grad():
                logw = mod(x, src_mask, sigma=sigma, inference=True)
        finally:
            torch.randn = orig
        assert len(draws) == 1 and tuple(draws[0].shape) == (bsz, 2, tp)
        # model.py:302-309 (duration_stochastic branch)
        dur = torch.ceil(torch.exp(logw + 1e-9)).masked_fill(logw == 0, 0)
        dur = torch.clamp(dur, min=0).int()
        for i in range(len(dur)):
            if dur[i][~src_mask[i]].sum() <= (~src_mask[i]).sum() // 2:
                dur[i][~src_mask[i]] = 1
        cases.append({"x": x, "src_mask": src_mask, "sigma": sigma, "noise": draws[0], "logw": logw, "duration_rounded": dur})
        print(f"case B={bsz} Tp={tp}: logw range [{float(logw.min()):.3f}, {float(logw.max()):.3f}], durations {dur[0].tolist()[:12]}")
    out = os.path.join(ROOT, "tests", "golden", "sdp_small.pt")
    torch.save({"cfg": CFG, "seed": seed, "shapes": {k: tuple(v.shape) for k, v in sd.items()}, "cases": cases}, out)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
