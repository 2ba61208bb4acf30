"""Run a few C2 synthesis steps (the bench workload) for ncu: `python tools/profile_step.py [steps] [mode]`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from lightningfastspeech2_b200 import synthetic  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
mode = sys.argv[2] if len(sys.argv) > 2 else "fp32"
dev = torch.device("cuda", 0)
model, sd, hp = bench.build_model(dev)
model.set_compute_mode(mode)
batch = {k: v.to(dev) for k, v in synthetic.make_batch(bench.BATCH, bench.MIN_LEN, bench.MAX_LEN, seed=2).items()
         if k in ("phones", "speaker")}
with torch.no_grad():
    for _ in range(steps):
        r = model(batch, inference=True)
torch.cuda.synchronize()
print("mel", tuple(r["mel"].shape), "frames", int((~r["tgt_mask"]).sum()))
