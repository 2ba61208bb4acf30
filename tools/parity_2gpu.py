#!/usr/bin/env python
"""Multi-GPU correctness on hardware.  Launch on N GPUs of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/parity_2gpu.py > profiles/<round>_parity_2gpu.log

(1) synthesis: every rank synthesises its shard of one batch; rank 0 then synthesises EVERY shard itself and each
    rank's mel / masks / durations must be bit-equal to rank 0's result for the same sub-batch (no collective on the
    data path, same kernels, same padded shapes => same bits on every GPU).
(2) training: every rank runs forward + loss + backward on its shard, then the ONE all-reduce (sum) over the flat
    gradient buffer.  The reduced buffer must equal (a) the sum of the per-shard gradients rank 0 computes alone, to
    fp32 atomics noise, and (b) the sum of the ORACLE's per-shard autograd gradients within the parity tolerance;
    scaled by 1/N it is the gradient of the mean of the per-rank mean losses = Lightning-DDP semantics
    (reference: Trainer(strategy="ddp") around fastspeech2.py:786-797).
Exit code != 0 on any mismatch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from lightningfastspeech2_b200 import configs, sharding, synthetic  # noqa: E402
from lightningfastspeech2_b200.fastspeech2.fastspeech2 import FastSpeech2  # noqa: E402


def build(preset, seed, dev, train=False, mode="fp32"):
    kw = configs.PRESETS[preset]
    hp = configs.resolve(kw)
    st = {v: {"min": -3.0, "max": 3.0, "mean": 0.0, "std": 1.0} for v in hp["variances"]}
    model = FastSpeech2(stats=st, phone2id={f"p{i}": i for i in range(80)}, num_workers=0, **kw)
    sd = synthetic.fill_state_dict(model.state_dict(), seed=seed)
    model.load_state_dict(sd)
    hp["stats"] = st
    model = model.to(dev)
    return (model.train() if train else model.eval()).set_compute_mode(mode), sd, hp


def gather_bytes_to_rank0(t, src, rank, dev):
    """rank `src` sends tensor t (any dtype / shape) to rank 0 -> (uint8 payload, shape list) on rank 0"""
    if rank == src:
        meta = torch.tensor([t.dim()] + list(t.shape) + [0] * (4 - t.dim()), dtype=torch.int64, device=dev)
        dist.send(meta, 0)
        dist.send(t.contiguous().view(torch.uint8).flatten(), 0)
        return None
    meta = torch.zeros(5, dtype=torch.int64, device=dev)
    dist.recv(meta, src)
    return meta.tolist()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    report, ok = {"world": world}, True

    # ---------------- (1) synthesis: sharded == single GPU on identical sub-batches ----------------
    for preset, mode in (("C2", "fp32"), ("C2", "bf16")):
        model, sd, hp = build(preset, 0, dev, mode=mode)
        full = synthetic.make_batch(16, 24, 200, seed=70)
        mine = sharding.shard_batch(full, rank, world)
        with torch.no_grad():
            r = model(mine, inference=True)
        keys = ("mel", "tgt_mask", "duration_rounded", "duration_prediction", "variances_pitch")
        same = True
        for src in range(world):
            r0 = None
            if rank == 0:
                with torch.no_grad():
                    r0 = model(sharding.shard_batch(full, src, world), inference=True)
            for k in keys:
                if src == 0:
                    if rank == 0:
                        same &= bool(torch.equal(r0[k], r[k]))
                elif rank == src:
                    gather_bytes_to_rank0(r[k], src, rank, dev)
                elif rank == 0:
                    meta = gather_bytes_to_rank0(None, src, rank, dev)
                    shape = meta[1:1 + meta[0]]
                    nbytes = r0[k].element_size()
                    for s_ in shape:
                        nbytes *= s_
                    buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                    dist.recv(buf, src)
                    same &= shape == list(r0[k].shape) and bool(
                        torch.equal(buf, r0[k].contiguous().view(torch.uint8).flatten()))
        if rank == 0:
            report[f"synthesis_{preset}_{mode}_bit_equal_to_single_gpu"] = same
            ok &= same
        dist.barrier()

    # ---------------- (2) training: all-reduced flat gradient == sum of per-shard gradients ----------------
    from oracle import fs2_oracle as O

    # simt = exact fp32 kernels: the sharp check.  fp32 mode (split-bf16 tensor cores): per-tensor relative L2 of the sum
    # of two small shards against the oracle; single ReLU-kink flips dominate the small tensors at these batch sizes
    for preset, mode, tol in (("SMALL_TRAIN", "simt", 1e-3), ("SMALL_TRAIN", "fp32", 1e-2)):
        model, sd, hp = build(preset, 1, dev, train=True, mode=mode)
        model.log_losses = False
        full = synthetic.add_train_targets(synthetic.make_batch(8, 6, 30, seed=71), hp["variances"], seed=71)
        flat_p, flat_g = model.flatten_parameters()

        def shard_grads(src):
            sub = sharding.shard_batch(full, src, world, max_frames=configs.max_frames(hp))
            flat_g.zero_()
            loss = model.training_step(sub, 0)
            loss.backward()
            torch.cuda.synchronize()
            return sub, float(loss.detach()), flat_g.clone()

        _, _, local_g = shard_grads(rank)
        flat_g.copy_(local_g)
        n = model.allreduce_gradients()
        reduced = flat_g.clone()
        if rank == 0:
            total = torch.zeros_like(local_g)
            ototal, losses = {}, []
            for src in range(world):
                s, l, g = shard_grads(src)
                total += g
                ol, og = O.gradients(sd, hp, s)
                losses.append((l, float(ol["total"])))
                for k, v in og.items():
                    ototal[k] = ototal.get(k, 0) + v
            scale = float(total.abs().max())
            e_self = float((reduced - total).abs().max()) / scale
            flat_g.copy_(reduced)  # compare per tensor through the parameters' gradient views
            worst = 0.0
            floor = 1e-2 * max(float(v.abs().max()) for v in ototal.values())
            for k, p in model.named_parameters():
                if k in ototal:
                    d = p.grad.cpu() - ototal[k]
                    worst = max(worst, float(d.norm()) / max(float(ototal[k].norm()), floor * d.numel() ** 0.5 * 0.1))
            good = (n == world and e_self < 1e-5 and worst < tol
                    and all(abs(a - b) < 1e-4 * max(1.0, abs(b)) for a, b in losses))
            report[f"train_{preset}_{mode}"] = {"world_returned": n, "allreduce_vs_own_sum_rel": e_self,
                                                "allreduce_vs_oracle_sum_worst_rel_l2": worst,
                                                "per_shard_loss_cuda_vs_oracle": losses, "ok": good}
            ok &= good
        dist.barrier()

    if rank == 0:
        report["ok"] = bool(ok)
        print(json.dumps(report, indent=1))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()
