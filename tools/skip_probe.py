"""Times model.skip_pad_rows on the bench batch: repeated resident loops, host time per step, per-kernel events."""
import sys, time
import torch
sys.path.insert(0, ".")
import bench
from lightningfastspeech2_b200 import ops, synthetic, _lib

dev = torch.device("cuda", 0)
model, sd, hp = bench.build_model(dev)
b = synthetic.make_batch(bench.BATCH, bench.MIN_LEN, bench.MAX_LEN, seed=2)
res = {k: v.to(dev) for k, v in b.items() if k in ("phones", "speaker")}

def loop(n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        with torch.no_grad():
            out = model(res, inference=True)
    e1.record()
    host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, host / n * 1e3

for skip in (False, True, False, True):
    model.skip_pad_rows = skip
    for r in range(3):
        print("skip", skip, "round", r, "gpu ms/step %.3f host ms/step %.3f" % loop(10), flush=True)
model.skip_pad_rows = True
ops.PROFILE = {}
with torch.no_grad():
    model(res, inference=True)
prof = ops.collect_profile()
ops.PROFILE = None
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{k:30s} {v['launches']:3d} {v['ms']:.3f}")
print("sum", sum(v["ms"] for v in prof.values()))

