"""A few HiFi-GAN generator calls on a ragged batch (for ncu / CUDA-event tables): python tools/profile_vocoder.py [utts] [--table]"""
import contextlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lightningfastspeech2_b200 import hifigan, ops, synthetic  # noqa: E402
from oracle import hifigan_oracle as HO  # noqa: E402

nutt = int([a for a in sys.argv[1:] if not a.startswith("--")][0]) if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else 8
dev = torch.device("cuda", 0)
gen = hifigan.Generator(hifigan.AttrDict(HO.CONFIG))
with contextlib.redirect_stdout(sys.stderr):
    gen.remove_weight_norm()
gen.load_state_dict(synthetic.hifigan_state_dict(HO.CONFIG, seed=3))
gen = gen.eval().to(dev)
g = torch.Generator().manual_seed(0)
lens = torch.randint(600, 2100, (nutt,), generator=g)
x = torch.randn(nutt, 80, int(lens.max()), generator=g).to(dev)
with torch.no_grad():
    for _ in range(2):
        w = gen(x, lens)
    torch.cuda.synchronize()
    if "--table" in sys.argv:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            gen(x, lens)
        e1.record()
        torch.cuda.synchronize()
        print(f"wall per call: {e0.elapsed_time(e1) / 3:.2f} ms for {int(lens.sum())} valid frames ({nutt} x {int(lens.max())} padded)")
        ops.PROFILE = {}
        gen(x, lens)
        prof = ops.collect_profile()
        ops.PROFILE = None
        tot = sum(v["ms"] for v in prof.values())
        print(f"sum of kernel times: {tot:.2f} ms over {sum(v['launches'] for v in prof.values())} launches")
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
            gbs = v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0
            tfs = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0
            print(f"{k:32s} {v['launches']:4d} launches {v['ms']:8.3f} ms {100 * v['ms'] / tot:5.1f}%  {gbs:8.0f} GB/s {tfs:7.1f} TF/s")
print("wav", tuple(w.shape))
