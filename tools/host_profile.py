"""Host-side cost of one synthesis step (python tools/host_profile.py [buckets]): cProfile of the launch path."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from lightningfastspeech2_b200 import synthetic  # noqa: E402

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
model, sd, hp = bench.build_model(dev)
model.length_buckets = nb
batch = {k: v.to(dev) for k, v in synthetic.make_batch(bench.BATCH, bench.MIN_LEN, bench.MAX_LEN, seed=2).items()
         if k in ("phones", "speaker")}
with torch.no_grad():
    for _ in range(3):
        model(batch, inference=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        model(batch, inference=True)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"host issue time per step {1e3 * (t1 - t0) / 5:.2f} ms, with sync {1e3 * (t2 - t0) / 5:.2f} ms")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        model(batch, inference=True)
    pr.disable()
    torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
