"""Depthwise conv at the C2 decoder's launch size, per kernel size: `LFS2_DWCONV_TMA=0|1|2 python tools/dwconv_ab.py`
(0 = per-thread-load kernel, 1 = TMA ring of 2 stages x 2 CTAs per SM, 2 = 4 stages x 1 CTA per SM)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lightningfastspeech2_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
b, t, d = 64, 2635, 256
g = torch.Generator().manual_seed(0)
xp = ops.split_bf16(torch.randn(b, t, d, generator=g).to(dev))
bias = torch.zeros(d, device=dev)
print("LFS2_DWCONV_TMA =", os.environ.get("LFS2_DWCONV_TMA", "(default)"))
for ks in (9, 13, 17, 21, 25):
    wt = (torch.randn(ks, d, generator=g) * 0.1).to(dev)
    for out, nbytes in (("f16", 6.0), ("planes", 8.0)):
        for _ in range(3):
            ops.dwconv1d_planes(xp, wt, bias, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.dwconv1d_planes(xp, wt, bias, out=out)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f"k = {ks:2d} out = {out:6s}: {us:7.1f} us  {nbytes * b * t * d / us / 1e3:7.0f} GB/s")
