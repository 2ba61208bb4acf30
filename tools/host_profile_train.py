"""Host-side cost of one C4 train step on a small shard (what a rank sees at N = 8: 8 utterances per GPU), where the
Python/ctypes launch path, not the GPU, sets the step time: `python tools/host_profile_train.py [utterances]`."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from lightningfastspeech2_b200 import _lib  # noqa: E402

nutt = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
model, sd, hp = bench.build_train_model(dev)
model.set_compute_mode("fp32")
model.log_losses = False
full = bench.train_batch(hp, 0, bench.TRAIN_BATCH // nutt)       # rank 0's shard of a world of 64 / nutt ranks
batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in full.items()}
print("shard:", tuple(batch["phones"].shape), "mel", tuple(batch["mel"].shape))
(opt,), (sch,) = model.configure_optimizers()
bench.run_train_steps(model, batch, opt, sch["scheduler"], 3, 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
c0 = _lib.CALLS
e0.record()
t0 = time.perf_counter()
bench.run_train_steps(model, batch, opt, sch["scheduler"], 5, 1)
t1 = time.perf_counter()
e1.record()
torch.cuda.synchronize()
print(f"device time per step {e0.elapsed_time(e1) / 5:.2f} ms, host issue time per step {1e3 * (t1 - t0) / 5:.2f} ms, "
      f"{(_lib.CALLS - c0) // 5} launches per step")
pr = cProfile.Profile()
pr.enable()
bench.run_train_steps(model, batch, opt, sch["scheduler"], 3, 1)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
