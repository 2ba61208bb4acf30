"""How far the single-pass bf16 train step is from the fp64 gradients (C4_P0 = the timed 76 M configuration with
dropout off, 4 utterances): per-tensor relative L2 error and cosine, next to the split-bf16 "fp32" mode.
`python tools/train_bf16_check.py`"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lightningfastspeech2_b200 import configs, synthetic  # noqa: E402
from lightningfastspeech2_b200.fastspeech2.fastspeech2 import FastSpeech2  # noqa: E402
from oracle import fs2_oracle as O  # noqa: E402

preset = sys.argv[1] if len(sys.argv) > 1 else "C4_P0"
kw = configs.PRESETS[preset]
hp = configs.resolve(kw)
st = {v: {"min": -3.0, "max": 3.0, "mean": 0.0, "std": 1.0} for v in hp["variances"]}
model = FastSpeech2(stats=st, phone2id={f"p{i}": i for i in range(80)}, num_workers=0, **kw)
sd = synthetic.fill_state_dict(model.state_dict(), seed=11)
model.load_state_dict(sd, strict=True)
hp["stats"] = st
model = model.to("cuda").train()
model.log_losses = False
batch = synthetic.add_train_targets(synthetic.make_batch(4, 24, 64, seed=11), hp["variances"], seed=11)
losses, g64 = O.gradients(sd, hp, batch, dtype=torch.float64)
print("oracle fp64 losses", {k: float(v) for k, v in losses.items()})
for mode in ("fp32", "bf16"):
    model.set_compute_mode(mode)
    model.zero_grad(set_to_none=True)
    model.training_step(batch, 0).backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.double().cpu() for k, p in model.named_parameters() if p.grad is not None}
    rows = []
    for k, w in g64.items():
        g = grads[k]
        rel = float((g - w).norm()) / max(float(w.norm()), 1e-30)
        cos = float((g * w).sum()) / max(float(g.norm()) * float(w.norm()), 1e-30)
        rows.append((rel, cos, k, float(w.norm())))
    rows.sort(reverse=True)
    flat_g = torch.cat([grads[k].flatten() for k in g64])
    flat_w = torch.cat([g64[k].flatten() for k in g64])
    print(f"[{mode}] loss {model.loss.last_buffer.tolist()[-1]:.6f}; whole gradient: rel L2 "
          f"{float((flat_g - flat_w).norm() / flat_w.norm()):.3e}, cosine {float((flat_g * flat_w).sum() / (flat_g.norm() * flat_w.norm())):.6f}")
    print(f"[{mode}] median per-tensor rel L2 {sorted(r[0] for r in rows)[len(rows) // 2]:.3e}; worst five:")
    for rel, cos, k, n in rows[:5]:
        print(f"    {k:60s} rel L2 {rel:.3e} cosine {cos:.6f} |g| {n:.3e}")
