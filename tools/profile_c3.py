"""CUDA-event kernel table of one C3 (76M, bf16) synthesis step: python tools/profile_c3.py [mode] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from lightningfastspeech2_b200 import ops, synthetic  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
bsz = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda", 0)
model, _, _ = bench.build_model(dev, preset="C3")
model.set_compute_mode(mode)
batch = {k: v.to(dev) for k, v in synthetic.make_batch(bsz, bench.MIN_LEN, bench.MAX_LEN, seed=200).items()
         if k in ("phones", "speaker")}
with torch.no_grad():
    for _ in range(3):
        r = model(batch, inference=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        model(batch, inference=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"wall per step: {e0.elapsed_time(e1) / 3:.2f} ms, mel {tuple(r['mel'].shape)}")
    ops.PROFILE = {}
    model(batch, inference=True)
    prof = ops.collect_profile()
    ops.PROFILE = None
tot = sum(v["ms"] for v in prof.values())
print(f"sum of kernel times: {tot:.2f} ms over {sum(v['launches'] for v in prof.values())} launches")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    gbs = v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0
    tfs = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0
    print(f"{k:32s} {v['launches']:4d} launches {v['ms']:8.3f} ms {100 * v['ms'] / tot:5.1f}%  {gbs:8.0f} GB/s {tfs:7.1f} TF/s")
