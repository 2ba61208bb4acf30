#!/usr/bin/env python
"""Golden vectors of the FastDiff variance adaptor from the UNMODIFIED reference (authoring container only):

    python oracle/make_goldens_fastdiff.py      # writes tests/golden/fastdiff_adaptor.pt

The reference's FastDiffVarianceAdaptor (litfass/fastspeech2/fastdiff_variances.py, imported through oracle/ref_shim.py)
gets seeded weights and runs on CPU (a) in inference mode (N = 4 reverse steps) and (b) teacher-forced, with every random
draw RECORDED: std_normal (both the module's and util.py's binding), torch.randint (diffusion steps) and torch.rand (the
duration jitter) are wrapped for the duration of the call.  The golden stores inputs, draws and every result tensor; the
weights are regenerated from the seed wherever it is replayed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lightningfastspeech2_b200 import synthetic  # noqa: E402
from oracle import ref_shim  # noqa: E402

CFG = dict(variances=["pitch", "energy"], variance_nlayers=[2, 2], variance_kernel_size=[3, 3], variance_dropout=[0.0, 0.0],
           variance_filter_size=256, variance_nbins=32, variance_depthwise_conv=True, duration_nlayers=2,
           duration_kernel_size=3, duration_dropout=0.0, duration_filter_size=256, duration_depthwise_conv=True,
           encoder_hidden=256, max_length=32 * 22050 / 256)


def main():
    if hasattr(ref_shim, "install"):
        ref_shim.install()
    import litfass.fastspeech2.fastdiff_variances as fv
    import litfass.third_party.fastdiff.module.util as util

    stats = {v: {"min": -3.0, "max": 3.0, "mean": 0.0, "std": 1.0} for v in CFG["variances"]}
    ada = fv.FastDiffVarianceAdaptor(stats, CFG["variances"], CFG["variance_nlayers"], CFG["variance_kernel_size"],
                                     CFG["variance_dropout"], CFG["variance_filter_size"], CFG["variance_nbins"],
                                     CFG["variance_depthwise_conv"], CFG["duration_nlayers"], CFG["duration_kernel_size"],
                                     CFG["duration_dropout"], CFG["duration_filter_size"], CFG["duration_depthwise_conv"],
                                     CFG["encoder_hidden"], CFG["max_length"])
    seed = 5
    sd = synthetic.fill_state_dict(ada.state_dict(), seed=seed)
    ada.load_state_dict(sd)
    ada.eval()

    g = torch.Generator().manual_seed(17)
    bsz, tp = 2, 11
    lens = [11, 7]
    x = torch.randn(bsz, tp, CFG["encoder_hidden"], generator=g)
    src_mask = torch.arange(tp)[None, :] >= torch.tensor(lens)[:, None]

    draws = {"noise": [], "randint": [], "rand": []}
    rng = torch.Generator().manual_seed(23)

    def rec_normal(size, device="cpu"):
        z = torch.normal(0, 1, size=tuple(size), generator=rng)
        draws["noise"].append(z.clone())
        return z

    orig = (fv.std_normal, util.std_normal, torch.randint, torch.rand)
    # the reference's helper defaults to device="cuda:0" and the module never passes one (fastdiff_variances.py:195):
    # bind the same function with device="cpu" (no arithmetic changes)
    _embed = util.calc_diffusion_step_embedding
    fv.calc_diffusion_step_embedding = lambda ts, dim: _embed(ts, dim, device="cpu")

    def rec_randint(high, size=None, **kw):
        t = orig[2](high, size=size, generator=rng)
        draws["randint"].append(t.clone())
        return t

    def rec_rand(size=None, **kw):
        t = orig[3](tuple(size), generator=rng)
        draws["rand"].append(t.clone())
        return t

    def run(targets, inference):
        for k in draws:
            draws[k] = []
        fv.std_normal = util.std_normal = rec_normal
        torch.randint, torch.rand = rec_randint, rec_rand
        try:
            with torch.no_grad():
                r = ada(x.clone(), src_mask, targets, inference=inference)
        finally:
            fv.std_normal, util.std_normal, torch.randint, torch.rand = orig
        return {k: (v.clone() if torch.is_tensor(v) else v) for k, v in r.items()}, {k: list(v) for k, v in draws.items()}

    infer, infer_draws = run({}, True)
    # teacher-forced: durations given, frame-level targets at the regulated (64-padded) length
    dur = torch.randint(1, 6, (bsz, tp), generator=g) * (~src_mask)
    tm = int(-(-int(dur.sum(1).max()) // 64) * 64)
    targets = {"duration": dur}
    for v in CFG["variances"]:
        targets[f"variances_{v}"] = torch.randn(bsz, tm, generator=g)
    train, train_draws = run(targets, False)

    golden = {"seed": seed, "cfg": dict(CFG, stats=stats), "state_dict_keys": sorted(sd), "x": x, "src_mask": src_mask,
              "inference": {"out": infer, "noise": infer_draws["noise"]},
              "teacher_forced": {"targets": targets, "out": train, "noise": train_draws["noise"],
                                 "steps": [t.reshape(-1) for t in train_draws["randint"]], "jitter": train_draws["rand"][0]}}
    path = os.path.join(ROOT, "tests", "golden", "fastdiff_adaptor.pt")
    torch.save(golden, path)
    print("wrote", path, os.path.getsize(path), "bytes; inference mel-side x", tuple(infer["x"].shape), "durations",
          infer["duration_rounded"].tolist(), "noise draws", len(infer_draws["noise"]), "/", len(train_draws["noise"]))


if __name__ == "__main__":
    main()
