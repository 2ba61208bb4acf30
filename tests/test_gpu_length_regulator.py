"""LengthRegulator on the GPU: bit-exact against the reference goldens and the oracle
(random, ragged, zero, truncated, int32/int64, bf16/fp32), plus size-independent
properties at the BASELINE config-5 stress size (B=512, Tp=400, d=256, dur ~ U{0..10})."""
import os

import numpy as np
import pytest
import torch

from lightningfastspeech2_b200 import ops
from oracle import fs2_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _bits(t):
    return t.contiguous().view(torch.int16 if t.element_size() == 2 else torch.int32)


def test_goldens_bit_exact(golden_dir):
    cases = torch.load(os.path.join(golden_dir, "length_regulator.pt"), weights_only=False)
    for c in cases:
        out, mask = ops.length_regulate(c["x"].to(DEV), c["dur"].to(DEV), c["max_length"])
        assert out.dtype == c["out"].dtype and tuple(out.shape) == tuple(c["out"].shape), c["tag"]
        assert torch.equal(mask.cpu(), c["mask"]), c["tag"]
        assert torch.equal(_bits(out.cpu()), _bits(c["out"])), c["tag"]


@pytest.mark.parametrize("seed", range(6))
def test_random_vs_oracle(seed):
    g = np.random.default_rng(seed)
    b, tp = int(g.integers(1, 9)), int(g.integers(1, 70))
    d = int(g.choice([4, 8, 64, 256]))
    hi = int(g.choice([1, 3, 10, 40]))
    dt = [torch.int32, torch.int64][seed % 2]
    x = torch.from_numpy(g.standard_normal((b, tp, d)).astype(np.float32))
    dur = torch.from_numpy(g.integers(0, hi + 1, (b, tp))).to(dt)
    cap = float(g.choice([2756.25, 17.5, 100.0]))
    ro, rm = O.length_regulator(x, dur, cap)
    out, mask = ops.length_regulate(x.to(DEV), dur.to(DEV), cap)
    assert torch.equal(mask.cpu(), rm)
    assert torch.equal(_bits(out.cpu()), _bits(ro))


def test_all_zero_durations_give_empty_output():
    x = torch.randn(3, 5, 8)
    out, mask = ops.length_regulate(x.to(DEV), torch.zeros(3, 5, dtype=torch.int64, device=DEV), 2756.25)
    assert tuple(out.shape) == (3, 0, 8) and tuple(mask.shape) == (3, 0)


@pytest.mark.parametrize("dt", [torch.int32, torch.int64])
def test_config5_stress_properties(dt):
    g = np.random.default_rng(5)
    b, tp, d = 512, 400, 256
    x = torch.from_numpy(g.standard_normal((b, tp, d)).astype(np.float32)).to(DEV)
    dur = torch.from_numpy(g.integers(0, 11, (b, tp))).to(dt).to(DEV)
    out, mask = ops.length_regulate(x, dur, 2756.25)
    lengths = dur.sum(1)
    L = min(int(lengths.max()), 2756)
    assert tuple(out.shape) == (b, L, d)
    # mask == !(t < len)
    t = torch.arange(L, device=DEV)
    assert torch.equal(mask, ~(t[None] < lengths[:, None]))
    # PAD frames are +0.0 bit patterns
    assert (out[mask].view(torch.int32) == 0).all()
    # segment-sum round trip: summing frames per phone == dur * x (exact in int index space:
    # check via a one-hot channel: frame t must carry the row of phone idx[t])
    idx_ref = torch.repeat_interleave(torch.arange(tp, device=DEV).repeat(b), dur.flatten().long())
    rows = x.reshape(b * tp, d)[torch.repeat_interleave(torch.arange(b * tp, device=DEV), dur.flatten().long())]
    valid = ~mask
    # frames are ordered (b, t): the valid ones in order are exactly `rows` (cut at L per utterance)
    keep = torch.ones(rows.shape[0], dtype=torch.bool, device=DEV)
    if int(lengths.max()) > L:
        pos = torch.cat([torch.arange(int(n), device=DEV) for n in lengths.tolist()])
        keep = pos < L
    assert torch.equal(out[valid].view(torch.int32), rows[keep].view(torch.int32))
    # a 16-utterance slice against the CPU oracle, bit for bit
    ro, rm = O.length_regulator(x[:16].cpu(), dur[:16].cpu(), 2756.25)
    l16 = ro.shape[1]
    assert torch.equal(out[:16, :l16].cpu().view(torch.int32), ro.view(torch.int32))
    assert (out[:16, l16:].view(torch.int32) == 0).all()
