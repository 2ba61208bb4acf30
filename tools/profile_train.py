"""Run a few C4 train steps (bench.py's "train" workload) -- for ncu, or stand-alone to print the
CUDA-event table of one step: `python tools/profile_train.py [steps] [mode] [preset] [--table] [--buckets=n] [--world=w]`
(--buckets: model.train_length_buckets; --world: time rank 0's shard of a w-rank job on this one GPU; the host's issue
time per step is printed beside the device time)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from lightningfastspeech2_b200 import ops  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
steps = int(args[0]) if len(args) > 0 else 2
mode = args[1] if len(args) > 1 else "fp32"
preset = args[2] if len(args) > 2 else bench.TRAIN_PRESET
dev = torch.device("cuda", 0)
model, sd, hp = bench.build_train_model(dev, preset)
model.set_compute_mode(mode)
model.log_losses = False
for a in sys.argv[1:]:
    if a.startswith("--buckets="):
        model.train_length_buckets = int(a.split("=")[1])
world = 1
for a in sys.argv[1:]:
    if a.startswith("--world="):
        world = int(a.split("=")[1])
batch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in bench.train_batch(hp, 0, world).items()}
(opt,), (sch,) = model.configure_optimizers()
bench.run_train_steps(model, batch, opt, sch["scheduler"], steps, 1)
torch.cuda.synchronize()
if "--table" in sys.argv:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    bench.run_train_steps(model, batch, opt, sch["scheduler"], 3, 1)
    t1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    print(f"train_length_buckets = {model.train_length_buckets}, mode {mode}")
    print(f"wall per step: {e0.elapsed_time(e1) / 3:.2f} ms (host issue time per step: {1e3 * (t1 - t0) / 3:.2f} ms)")
    ops.PROFILE = {}
    bench.run_train_steps(model, batch, opt, sch["scheduler"], 1, 1)
    prof = ops.collect_profile()
    ops.PROFILE = None
    tot = sum(v["ms"] for v in prof.values())
    print(f"sum of kernel times: {tot:.2f} ms over {sum(v['launches'] for v in prof.values())} launches")
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        gbs = v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0
        tfs = v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0
        print(f"{k:32s} {v['launches']:4d} launches {v['ms']:8.3f} ms {100 * v['ms'] / tot:5.1f}%  {gbs:8.0f} GB/s {tfs:7.1f} TF/s")
print("loss", model.loss.last_buffer.tolist())
