"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference's mel-generation path.

This file is the parity oracle for the CUDA path.  It must never be imported by the
product package (lightningfastspeech2_b200/); only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker or
as the timed CPU baseline.

What it restates (reference = MiniXC/LightningFastSpeech2 at /root/reference):
  FastSpeech2.forward                       litfass/fastspeech2/fastspeech2.py:636-736
  PositionalEncoding                        litfass/fastspeech2/model.py:38-55
  ConformerEncoderLayer (FFTBlock)          model.py:67-122  (+ torch MHA semantics)
  SpeakerEmbedding                          model.py:125-143
  VarianceAdaptor                           model.py:249-341
  LengthRegulator                           model.py:349-370
  VarianceEncoder (non-CWT branch)          model.py:409-441
  VariancePredictor/VarianceConvolutionLayer model.py:482-561
  FastSpeech2Loss (masked L1/MSE + total)   litfass/fastspeech2/loss.py:57-81,83-213
  NoamLR                                    litfass/fastspeech2/noam.py:20-25

All floating-point arithmetic of the reference lives in PyTorch (no torch pin in the
reference's pyproject.toml; torch 2.11.0 is what is installed), so this port calls the
same torch.nn.functional CPU primitives (conv1d / linear / layer_norm / softmax) that
the reference's nn.Modules dispatch to, which also makes it a fair CPU baseline.

PARITY PINNING: the reference has no tests or golden vectors of its own.
This oracle is pinned against outputs of the reference itself, run in the authoring
container under oracle/ref_shim.py, on seeded weights/inputs: oracle/make_goldens.py
wrote tests/golden/*.pt and tests/test_oracle_golden.py replays them everywhere.

Parameters come in as a plain {name: tensor} dict with the reference's state_dict key
names; hyper-parameters as a dict of the reference's constructor kwargs.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

LN_EPS = 1e-5  # nn.LayerNorm default, used by every LayerNorm on the path


def positional_table(sd, dtype):
    return sd["positional_encoding.pe"].to(dtype)


def speaker_term(sd, speaker, dtype):
    """relu(Linear(256,d)(dvec)) -- reference model.py:137-143 (broadcast over T by caller)."""
    w = sd["speaker_embedding.projection.weight"].to(dtype)
    b = sd["speaker_embedding.projection.bias"].to(dtype)
    return torch.relu(F.linear(speaker.to(dtype), w, b))


def front_end(sd, phones, speaker, dtype):
    """fastspeech2.py:651-660: embedding (padding_idx row is whatever the table holds),
    + PE[:Tp], + speaker term.  Dropout is identity (eval / p=0 parity runs)."""
    emb = sd["phone_embedding.weight"].to(dtype)
    x = emb[phones]
    x = x + positional_table(sd, dtype)[:, : phones.shape[1], :]
    spk = speaker_term(sd, speaker, dtype)
    return x + spk[:, None, :], spk


def self_attention(x, kpm, w_in, b_in, w_out, b_out, nhead):
    """torch _sa_block -> nn.MultiheadAttention semantics used at model.py:111-114:
    packed in-proj rows [Wq;Wk;Wv], heads = contiguous d/h column blocks, q scaled by
    (d/h)^-1/2 before QK^T, -inf on PAD keys, softmax over keys, then out-proj."""
    bsz, t, d = x.shape
    dh = d // nhead
    qkv = F.linear(x, w_in, b_in)
    q, k, v = qkv.split(d, dim=-1)

    def heads(z):
        return z.reshape(bsz, t, nhead, dh).permute(0, 2, 1, 3)

    q, k, v = heads(q), heads(k), heads(v)
    s = torch.matmul(q * (dh ** -0.5), k.transpose(-1, -2))
    if kpm is not None:
        s = s.masked_fill(kpm[:, None, None, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    a = torch.matmul(p, v).permute(0, 2, 1, 3).reshape(bsz, t, d)
    return F.linear(a, w_out, b_out)


def conv_ffn(x, sd, pre, depthwise, dtype):
    """_ff_block, model.py:118-122 with the conv definitions at model.py:73-106.
    Channels-first Conv1d over the *padded* time axis; zero "same" padding at the two
    ends of the padded tensor only (PAD rows inside the batch are real inputs)."""
    g = lambda n: sd[pre + n].to(dtype)
    xt = x.transpose(1, 2)
    d = xt.shape[1]
    if depthwise:
        k1 = sd[pre + "conv1.0.weight"].shape[-1]
        if k1 % 2 != 1:
            raise NotImplementedError("even kernel sizes ('same' pads asymmetrically)")
        u = F.conv1d(xt, g("conv1.0.weight"), g("conv1.0.bias"), padding=(k1 - 1) // 2, groups=d)
        v = torch.relu(F.conv1d(u, g("conv1.1.weight"), g("conv1.1.bias")))
        # conv2.0: groups = conv_in (model.py:90) on F channels => F/d in, F/d out per group
        w = F.conv1d(v, g("conv2.0.weight"), g("conv2.0.bias"), groups=d)
        y = F.conv1d(w, g("conv2.1.weight"), g("conv2.1.bias"))
    else:
        k1 = sd[pre + "conv1.weight"].shape[-1]
        k2 = sd[pre + "conv2.weight"].shape[-1]
        if k1 % 2 != 1 or k2 % 2 != 1:
            raise NotImplementedError("even kernel sizes")
        v = torch.relu(F.conv1d(xt, g("conv1.weight"), g("conv1.bias"), padding=(k1 - 1) // 2))
        y = F.conv1d(v, g("conv2.weight"), g("conv2.bias"), padding=(k2 - 1) // 2)
    return y.transpose(1, 2)


def fft_block(x, kpm, sd, pre, nhead, depthwise, dtype):
    """ConformerEncoderLayer.forward, post-norm branch (model.py:113-116)."""
    g = lambda n: sd[pre + n].to(dtype)
    d = x.shape[-1]
    a = self_attention(x, kpm, g("self_attn.in_proj_weight"), g("self_attn.in_proj_bias"),
                       g("self_attn.out_proj.weight"), g("self_attn.out_proj.bias"), nhead)
    x1 = F.layer_norm(x + a, (d,), g("norm1.weight"), g("norm1.bias"), LN_EPS)
    y = conv_ffn(x1, sd, pre, depthwise, dtype)
    return F.layer_norm(x1 + y, (d,), g("norm2.weight"), g("norm2.bias"), LN_EPS)


def variance_predictor(x, mask, sd, pre, nlayers, depthwise, dtype, return_hidden=False):
    """VariancePredictor.forward (model.py:510-522) over VarianceConvolutionLayer
    (model.py:524-561): per layer conv -> ReLU -> LayerNorm(filter) (-> dropout);
    then Linear(filter,1), squeeze, masked_fill(mask, 0)."""
    g = lambda n: sd[pre + n].to(dtype)
    z = x
    for l in range(nlayers):
        lp = f"layers.{l}.layers."
        zt = z.transpose(1, 2)
        if depthwise:
            k = sd[pre + lp + "0.module.0.weight"].shape[-1]
            c = zt.shape[1]
            u = F.conv1d(zt, g(lp + "0.module.0.weight"), g(lp + "0.module.0.bias"),
                         padding=(k - 1) // 2, groups=c)
            u = F.conv1d(u, g(lp + "0.module.1.weight"), g(lp + "0.module.1.bias"))
        else:
            k = sd[pre + lp + "0.module.weight"].shape[-1]
            u = F.conv1d(zt, g(lp + "0.module.weight"), g(lp + "0.module.bias"), padding=(k - 1) // 2)
        u = torch.relu(u).transpose(1, 2)
        z = F.layer_norm(u, (u.shape[-1],), g(lp + "2.weight"), g(lp + "2.bias"), LN_EPS)
    out = F.linear(z, g("linear.weight"), g("linear.bias")).squeeze(-1)
    if mask is not None:
        out = out.masked_fill(mask, 0)
    return (out, z) if return_hidden else out


def round_durations(log_dur, src_mask):
    """Inference duration rule, model.py:300-309: round-half-even(exp(p)-1), clamp>=0,
    int32; then per utterance, if sum over valid phones <= n_valid // 2, set every valid
    phone's duration to 1."""
    dur = torch.clamp(torch.round(torch.exp(log_dur) - 1), min=0).to(torch.int32)
    valid = ~src_mask
    nvalid = valid.sum(1)
    total = (dur * valid).sum(1)
    fix = total <= torch.div(nvalid, 2, rounding_mode="floor")
    dur = torch.where(fix[:, None] & valid, torch.ones_like(dur), dur)
    return dur


def length_regulator_indices(durations, max_length):
    """Integer core of LengthRegulator.forward (model.py:349-370), numpy int64.

    The reference repeats row p of utterance b durations[b,p] times, pads to the longest
    utterance with +0.0 and cuts at L = min(max_b len_b, int(max_length)).  Frame t of
    utterance b therefore reads phone idx[b,t] = #{p : cumsum_incl(dur[b])[p] <= t} when
    t < len_b.  Returns (idx (B,L) int64 [-1 on PAD frames], lengths (B,) int64, L)."""
    dur = np.asarray(durations).astype(np.int64)
    cum = np.cumsum(dur, axis=1)
    lengths = cum[:, -1] if dur.shape[1] else np.zeros(dur.shape[0], np.int64)
    L = int(min(int(lengths.max()) if lengths.size else 0, int(max_length)))
    idx = np.full((dur.shape[0], L), -1, dtype=np.int64)
    t = np.arange(L)
    for b in range(dur.shape[0]):
        n = int(min(lengths[b], L))
        idx[b, :n] = np.searchsorted(cum[b], t[:n], side="right")
    return idx, lengths, L


def length_regulator(x, durations, max_length):
    """Returns (out (B,L,d) same dtype as x with +0.0 on PAD frames, mask (B,L) bool True=PAD)."""
    idx, lengths, L = length_regulator_indices(durations.cpu().numpy(), max_length)
    idx_t = torch.from_numpy(idx)
    valid = idx_t >= 0
    gat = idx_t.clamp(min=0)
    out = torch.gather(x, 1, gat[:, :, None].expand(-1, -1, x.shape[-1]))
    out = torch.where(valid[:, :, None], out, torch.zeros((), dtype=x.dtype))
    mask = ~(torch.arange(L)[None, :] < torch.from_numpy(lengths)[:, None])
    return out, mask


def bucket_indices(values, bins):
    """torch.bucketize(values, bins) with right=False (model.py:422,436-438): number of
    boundaries strictly below the value."""
    return torch.bucketize(values, bins)


def variance_encoder(x, tgt, mask, sd, pre, nlayers, depthwise, mean, std, dtype, control=1.0, forced_idx=None):
    """VarianceEncoder.forward, non-CWT branch (model.py:409-441): prediction = predictor(x, mask); the embedding is
    looked up at bucketize(tgt * std + mean) when a target is given (:417-422), else at bucketize(prediction * std +
    mean) of the UNSCALED prediction, after which the returned prediction is multiplied by `control` (:434-438).
    -> (prediction, embedding, bucket index)"""
    pred = variance_predictor(x, mask, sd, pre + "predictor.", nlayers, depthwise, dtype)
    bins = sd[pre + "bins"].to(dtype)
    if forced_idx is not None:
        idx = forced_idx
    elif tgt is not None:
        idx = bucket_indices(tgt.to(dtype) * std + mean, bins)
    else:
        idx = bucket_indices(pred * std + mean, bins)
    if tgt is None:
        pred = pred * control
    return pred, sd[pre + "embedding.weight"].to(dtype)[idx], idx


def forward(sd, hp, batch, inference=False, dtype=torch.float32, force=None, control=None):
    """FastSpeech2.forward (fastspeech2.py:636-736) -> dict with the reference's result keys
    plus '_'-prefixed intermediates used by per-stage parity tests.

    force: optional {"duration_rounded": (B,Tp) int, "bucket_idx": {var: (B,Tm) int64}}
    to teacher-force the two discrete decision points."""
    force = force or {}
    control = control or {}
    phones = batch["phones"]
    src_mask = phones.eq(0)
    x, spk = front_end(sd, phones, batch["speaker"], dtype)
    res = {"_x0": x}

    for i in range(hp["encoder_layers"]):
        x = fft_block(x, src_mask, sd, f"encoder.layers.{i}.", hp["encoder_head"],
                      hp["encoder_depthwise_conv"], dtype)
    for prior in hp.get("priors", []):  # PriorEmbedding (model.py:146-164) added after the encoder (fastspeech2.py:687-692)
        pre = f"prior_embeddings.{prior}."
        vals = torch.as_tensor(batch[f"priors_{prior}"]).to(dtype)
        idx = torch.bucketize(vals, sd[pre + "bins"].to(dtype))
        x = x + torch.relu(sd[pre + "embedding.weight"].to(dtype)[idx])[:, None, :]
    res["_enc"] = x

    va = "variance_adaptor."
    log_dur = variance_predictor(x, src_mask, sd, va + "duration_predictor.", hp["duration_nlayers"],
                                 hp["duration_depthwise_conv"], dtype)
    def encode(i, var, x, mask, out_val):
        """VarianceEncoder.forward + the `x = x + out` of the adaptor (model.py:409-441, 284-294 / 315-333)"""
        pre = va + f"encoders.{var}."
        stats = hp.get("stats", {}).get(var, {"mean": 0.0, "std": 1.0})
        pred, emb, idx = variance_encoder(
            x, None if inference else batch[f"variances_{var}"], mask, sd, pre, hp["variance_nlayers"][i],
            hp["variance_depthwise_conv"], stats["mean"], stats["std"], dtype, control=control.get(var, 1.0),
            forced_idx=force["bucket_idx"][var] if "bucket_idx" in force and var in force["bucket_idx"] else None)
        res[f"variances_{var}"] = pred
        res[f"_bucket_{var}"] = idx
        return x + emb, (emb if out_val is None else out_val + emb)

    out_val = None
    for i, var in enumerate(hp["variances"]):  # phone-level variances act on the encoder output (model.py:277-294)
        if hp["variance_transforms"][i] == "cwt":
            raise NotImplementedError("cwt")
        if hp["variance_levels"][i] == "phone":
            x, out_val = encode(i, var, x, src_mask, out_val)
    if "duration_rounded" in force:
        dur = force["duration_rounded"]
    elif not inference:
        dur = batch["duration"]
    else:
        dur = round_durations(log_dur, src_mask)

    from_cfg = hp["max_length"] * hp["sampling_rate"] / hp["hop_length"]
    x, tgt_mask = length_regulator(x, dur, from_cfg)
    if out_val is not None:
        out_val, _ = length_regulator(out_val, dur, from_cfg)
    res["_lr"] = x

    for i, var in enumerate(hp["variances"]):
        if hp["variance_levels"][i] == "frame":
            x, out_val = encode(i, var, x, tgt_mask, out_val)
    res["_va"] = x

    x = x + positional_table(sd, dtype)[:, : x.shape[1], :]
    x = x + spk[:, None, :]
    for i in range(hp["decoder_layers"]):
        x = fft_block(x, tgt_mask, sd, f"decoder.layers.{i}.", hp["decoder_head"],
                      hp["decoder_depthwise_conv"], dtype)
    res["_dec"] = x
    res["mel"] = F.linear(x, sd["linear.weight"].to(dtype), sd["linear.bias"].to(dtype))
    res["duration_prediction"] = log_dur
    res["duration_rounded"] = dur
    res["src_mask"] = src_mask
    res["tgt_mask"] = tgt_mask
    if "fastdiff_linear.0.weight" in sd and out_val is not None:
        h = out_val + spk[:, None, :]
        for j in (0, 1):
            h = F.linear(h, sd[f"fastdiff_linear.{j}.weight"].to(dtype), sd[f"fastdiff_linear.{j}.bias"].to(dtype))
        res["fastdiff_var"] = h * 0.1
    return res


def loss(hp, result, batch):
    """FastSpeech2Loss.forward default branches (loss.py:83-213): masked MSE per variance,
    masked L1 mel, masked MSE log-duration, weighted total (weights fastspeech2.py:444-450)."""
    valid_src = ~result["src_mask"]
    valid_tgt = ~result["tgt_mask"]
    dt = result["mel"].dtype
    cap = int(hp["max_length"] * hp["sampling_rate"] / hp["hop_length"])
    out = {}
    for i, var in enumerate(hp["variances"]):
        if hp["variance_levels"][i] == "phone":  # loss.py:129-130: phone-level variances use the source mask
            tgt = batch[f"variances_{var}"].to(dt)
            out[var] = F.mse_loss(result[f"variances_{var}"][valid_src], tgt[valid_src])
        else:
            tgt = batch[f"variances_{var}"][:, :cap].to(dt)
            out[var] = F.mse_loss(result[f"variances_{var}"][valid_tgt], tgt[valid_tgt])
    m = valid_tgt[:, :, None].expand_as(result["mel"])
    out["mel"] = F.l1_loss(result["mel"][m], batch["mel"].to(dt)[m])
    out["duration"] = F.mse_loss(result["duration_prediction"][valid_src],
                                 torch.log(batch["duration"] + 1).to(dt)[valid_src])
    w = {"mel": hp["mel_loss_weight"], "duration": hp["duration_loss_weight"]}
    for i, var in enumerate(hp["variances"]):
        w[var] = hp["variance_loss_weights"][i]
    out["total"] = sum(v * w[k] for k, v in out.items())
    return out


def noam_scale(step, warmup):
    """NoamLR.get_lr scale factor (noam.py:20-25)."""
    s = max(1, step)
    return warmup ** 0.5 * min(s ** -0.5, s * warmup ** -1.5)


def gradients(sd, hp, batch, dtype=torch.float32):
    """Loss values and d(total)/d(parameter) of the teacher-forced train step (dropout off), by
    torch.autograd over this restatement -- what the reference's training_step + backward()
    computes (fastspeech2.py:786-797).  Returns ({loss name: float}, {param name: grad tensor})."""
    leaves = {}
    for k, v in sd.items():
        if k.endswith(".bins") or k.endswith("positional_encoding.pe") or not v.is_floating_point():
            leaves[k] = v
        else:
            leaves[k] = v.detach().to(dtype).clone().requires_grad_(True)
    b = {k: (v.to(dtype) if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in batch.items()}
    r = forward(leaves, hp, b, inference=False, dtype=dtype)
    ls = loss(hp, r, b)
    ls["total"].backward()
    grads = {k: v.grad for k, v in leaves.items() if torch.is_tensor(v) and v.requires_grad and v.grad is not None}
    if "phone_embedding.weight" in grads:
        grads["phone_embedding.weight"][0] = 0  # nn.Embedding(padding_idx=0): no gradient for the PAD row
    return {k: float(v.detach()) for k, v in ls.items()}, grads


def adamw_noam_step(param, grad, exp_avg, exp_avg_sq, step, base_lr, warmup, betas=(0.9, 0.98), eps=1e-8,
                    weight_decay=0.01):
    """One AdamW update with the Noam-scheduled learning rate, as configure_optimizers sets it up
    (fastspeech2.py:1166-1182; noam.py:20-25): the scheduler has been stepped `step - 1` times when
    optimizer step number `step` (1-based) runs.  Plain tensor arithmetic, torch.optim.AdamW's order.
    Returns (new_param, new_exp_avg, new_exp_avg_sq, lr)."""
    lr = base_lr * noam_scale(step - 1, warmup)
    b1, b2 = betas
    p = param * (1 - lr * weight_decay)
    m = exp_avg + (grad - exp_avg) * (1 - b1)
    v = exp_avg_sq * b2 + (1 - b2) * grad * grad
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * (m / denom)
    return p, m, v, lr
