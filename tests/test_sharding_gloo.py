"""Multi-process (world_size 2, gloo, CPU) checks of the data-parallel host logic:
snake sharding covers every utterance exactly once and balances the load; the one all-reduce
of the flat gradient buffer + the 1/world factor reproduces Lightning-DDP semantics (mean of the
per-rank losses) -- checked against the oracle's autograd gradients of each shard."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lightningfastspeech2_b200 import configs, sharding, synthetic
from oracle import fs2_oracle as O


def test_snake_shards_partition_and_balance():
    g = torch.Generator().manual_seed(0)
    lengths = torch.randint(32, 513, (64,), generator=g).tolist()
    for world in (1, 2, 4, 8):
        shards = sharding.snake_shards(lengths, world)
        flat = sorted(u for s in shards for u in s)
        assert flat == list(range(64))
        assert {len(s) for s in shards} == {64 // world}
        loads = [sum(lengths[u] for u in s) for s in shards]
        assert max(loads) - min(loads) <= 0.05 * max(loads) + 512
        sq = [sum(lengths[u] ** 2 for u in s) for s in shards]
        assert max(sq) <= 1.25 * min(sq)


def test_shard_batch_pads_to_own_maximum():
    batch = synthetic.make_batch(8, 5, 40, seed=3)
    batch = synthetic.add_train_targets(batch, ["pitch", "energy"], seed=3)
    seen = []
    for rank in range(2):
        sub = sharding.shard_batch(batch, rank, 2)
        assert sub["phones"].shape[0] == 4
        assert sub["phones"].shape[1] == int(sub["phones_lengths"].max())
        assert (sub["phones"] != 0).sum(1).tolist() == sub["phones_lengths"].tolist()
        assert sub["duration"].shape == sub["phones"].shape
        assert sub["mel"].shape[0] == 4 and sub["speaker"].shape == (4, 256)
        # frame-level targets are cut to the shard's own longest expanded length: what the shard's model returns
        tm = int(sub["duration"].sum(1).max())
        assert sub["mel"].shape[1] == tm
        assert all(sub[f"variances_{v}"].shape == (4, tm) for v in ("pitch", "energy"))
        seen += sub["phones_lengths"].tolist()
    assert sorted(seen) == sorted(batch["phones_lengths"].tolist())


def test_shard_batch_cuts_phone_level_variances_and_caps_frames():
    batch = synthetic.make_batch(6, 5, 30, seed=5)
    batch = synthetic.add_train_targets(batch, ["pitch", "energy"], seed=5, levels=["phone", "frame"])
    for rank in range(2):
        sub = sharding.shard_batch(batch, rank, 2, max_frames=50)
        tp = int(sub["phones_lengths"].max())
        tm = min(int(sub["duration"].sum(1).max()), 50)
        assert sub["variances_pitch"].shape == (3, tp)      # phone level: cut like `phones`
        assert sub["variances_energy"].shape == (3, tm)     # frame level: cut like `mel`, capped
        assert sub["mel"].shape[:2] == (3, tm)


def test_phone_lengths_is_last_valid_plus_one():
    phones = torch.tensor([[3, 0, 4, 0, 0], [1, 2, 3, 4, 5], [0, 0, 0, 0, 0]])
    assert sharding.phone_lengths(phones).tolist() == [3, 5, 0]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    kw = configs.PRESETS["TINY_DW"]
    hp = configs.resolve(dict(kw, encoder_dropout=0.0, decoder_dropout=0.0))
    hp["stats"] = {v: {"min": -3.0, "max": 3.0, "mean": 0.0, "std": 1.0} for v in hp["variances"]}
    shapes = torch.load(os.path.join(os.path.dirname(__file__), "golden", "tiny_dw_train.pt"), weights_only=False)["shapes"]
    sd = synthetic.fill_state_dict({k: v for k, v in shapes.items() if not k.startswith("fastdiff")}, seed=9)
    full = synthetic.add_train_targets(synthetic.make_batch(6, 5, 17, seed=9), hp["variances"], seed=9)
    # the collated FULL batch goes straight through shard_batch into model + loss (no re-slicing by hand)
    sub = sharding.shard_batch(full, rank, world, max_frames=configs.max_frames(hp))
    losses, grads = O.gradients(sd, hp, sub)
    names = sorted(grads)
    flat = torch.cat([grads[k].flatten() for k in names])
    local = flat.clone()
    n = sharding.allreduce_sum_(flat)
    assert n == world
    torch.save({"local": local, "reduced": flat, "loss": losses["total"], "names": names}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_world2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / f"r{i}.pt", weights_only=False) for i in range(world)]
    total = r[0]["local"] + r[1]["local"]
    # both ranks hold the identical sum; with grad_scale = 1/world it is the gradient of the mean of the
    # per-rank mean losses (Lightning DDP semantics)
    assert torch.equal(r[0]["reduced"], r[1]["reduced"])
    assert torch.allclose(r[0]["reduced"], total, rtol=0, atol=0)
    assert r[0]["names"] == r[1]["names"]
    assert abs(r[0]["loss"] - r[1]["loss"]) > 0  # different shards, different losses
