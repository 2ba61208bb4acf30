"""Utterance sharding of the path across GPUs.

Utterances are independent (LayerNorm is per token, attention per utterance, nothing in
FastSpeech2.forward crosses utterances), so N GPUs = N processes that each take a slice of
the batch: no data-path collective in synthesis; training adds exactly one all-reduce (sum)
over the flat fp32 gradient buffer, the 1/N average being folded into the fused AdamW kernel
(Lightning-DDP semantics: the mean of the per-rank mean losses).

The reference sorts by duration for batching (dataset/datasets.py:884-886, used at
fastspeech2.py:1309-1310); dealing the length-sorted utterances in snake order balances both
sum(T) and sum(T^2) (attention) across ranks.
"""
import torch


def snake_shards(lengths, world):
    """lengths: sequence of utterance lengths -> list (one per rank) of utterance indices."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    shards = [[] for _ in range(world)]
    for pos, u in enumerate(order):
        r = pos % (2 * world)
        shards[r if r < world else 2 * world - 1 - r].append(u)
    return shards


def phone_lengths(phones):
    """(B,Tp) phoneme ids, 0 = PAD -> last valid position + 1 per utterance (not the count of non-PAD ids)"""
    tp = phones.shape[1]
    return ((phones != 0) * torch.arange(1, tp + 1, device=phones.device)).amax(1) if tp else phones.new_zeros(len(phones))


def shard_batch(batch, rank, world, length_key="phones_lengths", max_frames=None):
    """The rank's slice of a collated batch dict (reference dataset/datasets.py:852-882 layout):
    every tensor whose first dimension is the batch is indexed; phoneme-level tensors (`phones`, `duration`,
    phone-level `variances_*`) are cut to the shard's own longest utterance and frame-level ones (`mel`, frame-level
    `variances_*`) to the shard's own longest expanded length sum(duration), capped by `max_frames` (the
    LengthRegulator's limit) -- each rank pads to ITS longest utterance, which is also the shape its model returns."""
    lengths = batch[length_key].tolist() if length_key in batch else phone_lengths(batch["phones"]).tolist()
    mine = snake_shards(lengths, world)[rank]
    bsz = batch["phones"].shape[0]
    tp_full = batch["phones"].shape[1]
    keep = max(int(lengths[u]) for u in mine) if mine else 0
    tm_full = batch["mel"].shape[1] if torch.is_tensor(batch.get("mel")) else None
    tm_keep = None
    if tm_full is not None and torch.is_tensor(batch.get("duration")) and mine:
        tm_keep = int(batch["duration"][mine][:, :keep].sum(1).max())
        if max_frames is not None:
            tm_keep = min(tm_keep, int(max_frames))
    out = {}
    for k, v in batch.items():
        if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == bsz:
            v = v[mine]
            phone_level = k in ("phones", "duration") or (k.startswith("variances_") and v.dim() == 2
                                                          and v.shape[1] == tp_full and v.shape[1] != tm_full)
            frame_level = tm_keep is not None and (k == "mel" or (k.startswith("variances_") and v.dim() == 2
                                                                  and v.shape[1] == tm_full))
            if v.dim() >= 2 and v.shape[1] == tp_full and phone_level:
                v = v[:, :keep]
            elif frame_level:
                v = v[:, :tm_keep]
            out[k] = v.contiguous()
        else:
            out[k] = v
    return out


def allreduce_sum_(flat, group=None):
    """In-place sum of one flat buffer over the data-parallel group (NCCL on GPUs, gloo in the CPU
    tests); returns the world size the caller divides by (FusedAdamW.grad_scale = 1 / world)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return 1
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return world
