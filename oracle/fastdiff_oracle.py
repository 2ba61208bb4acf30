"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's FastDiff variance adaptor
(litfass/fastspeech2/fastdiff_variances.py:8-341 with the helpers of third_party/fastdiff/module/util.py:158-228, 276-343),
on a plain {name: tensor} state dict, with every random draw INJECTED (noise list in the reference's draw order, diffusion
steps, duration jitter) so that the CUDA path can be compared on identical draws.

Only tests/ may import this file.  PINNED against the unmodified reference module run in the authoring container with its
draws recorded (tests/golden/fastdiff_adaptor*.pt, written by oracle/make_goldens_fastdiff.py; tests/test_oracle_fastdiff.py).
"""
import math

import torch
import torch.nn.functional as F

from oracle import fs2_oracle as O

SCHEDULES = {
    8: [6.689325005027058e-07, 1.0033881153503899e-05, 0.00015496854030061513, 0.002387222135439515,
        0.035597629845142365, 0.3681158423423767, 0.4735414385795593, 0.5],
    6: [1.7838445955931093e-06, 2.7984189728158526e-05, 0.00043231004383414984, 0.006634317338466644,
        0.09357017278671265, 0.6000000238418579],
    4: [3.2176e-04, 2.5743e-03, 2.5376e-02, 7.0414e-01],
    3: [9.0000e-05, 9.0000e-03, 6.0000e-01],
}


def hyperparams(beta_0=1e-6, beta_T=0.01, T=1000):
    """compute_hyperparams_given_schedule(torch.linspace(beta_0, beta_T, T)), util.py:276-302"""
    beta = torch.linspace(beta_0, beta_T, T)
    alpha = 1 - beta
    sigma = beta + 0
    for t in range(1, T):
        alpha[t] *= alpha[t - 1]
        sigma[t] *= (1 - alpha[t - 1]) / (1 - alpha[t])
    return {"T": T, "beta": beta, "alpha": torch.sqrt(alpha), "sigma": torch.sqrt(sigma)}


def step_embedding(ts, dim):
    """calc_diffusion_step_embedding, util.py:318-343: ts (B, 1) -> (B, dim)"""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half) * -e)
    e = ts * e
    return torch.cat((torch.sin(e), torch.cos(e)), 1)


def denoise(sd, pre, hp, xt, c, ts, mask=None):
    """FastDiffVariancePredictor.forward with a given step (fastdiff_variances.py:192-221): xt (B, L), c (B, L, d)
    channels-last, ts (B, 1) -> noise prediction (B, L)"""
    g = lambda n: sd[pre + n]
    emb = step_embedding(ts, g("fc_t1.weight").shape[1])
    emb = F.linear(emb, g("fc_t1.weight"), g("fc_t1.bias"))
    emb = emb * torch.sigmoid(emb)
    emb = F.linear(emb, g("fc_t2.weight"), g("fc_t2.bias"))
    emb = emb * torch.sigmoid(emb)
    ne = F.linear(emb, g("linear_noise.weight"), g("linear_noise.bias"))          # (B, d)
    x = F.linear(xt.unsqueeze(-1), g("linear_in.weight"), g("linear_in.bias"))     # (B, L, d)
    inp = x + c + ne[:, None, :]
    return O.variance_predictor(inp, mask, sd, pre, hp["nlayers"], hp["depthwise"], torch.float32)


def q_sample(x0, z, steps, hyper):
    """x_t = alpha_t x_0 + sqrt(1 - alpha_t^2) z   (:177-190); x0 (B, L), z (B, 1, L) as the reference draws it, steps (B)"""
    a = hyper["alpha"][steps][:, None]
    return a * x0 + (1 - a ** 2.0).sqrt() * z.reshape(x0.shape)


def predictor_inference(sd, pre, hp, c, n_steps, noise, hyper):
    """FastDiffVariancePredictor.inference -> sampling_given_noise_schedule (util.py:158-228), ddim=False"""
    schedule = torch.FloatTensor(SCHEDULES[n_steps]) if n_steps in SCHEDULES else (
        torch.linspace(0.000001, 0.01, 1000) if n_steps == 1000 else torch.linspace(0.0001, 0.02, 200))
    alpha = hyper["alpha"]
    n = len(schedule)
    beta_infer = schedule
    alpha_infer = 1 - beta_infer
    sigma_infer = beta_infer + 0
    for i in range(1, n):
        alpha_infer[i] *= alpha_infer[i - 1]
        sigma_infer[i] *= (1 - alpha_infer[i - 1]) / (1 - alpha_infer[i])
    alpha_infer, sigma_infer = torch.sqrt(alpha_infer), torch.sqrt(sigma_infer)
    steps = []
    for i in range(n):
        ai = alpha_infer[i]
        if ai < alpha[-1]:
            steps.append(len(alpha) - 1)
        elif ai > alpha[0]:
            steps.append(0)
        else:
            for t in range(len(alpha) - 1):
                if alpha[t + 1] <= ai <= alpha[t]:
                    steps.append(t + ((alpha[t] - ai) / (alpha[t] - alpha[t + 1])).item())
                    break
    steps = torch.FloatTensor(steps)
    bsz, length = c.shape[0], c.shape[1]
    x = noise.pop(0)
    for i in range(len(steps) - 1, -1, -1):
        ts = steps[i] * torch.ones((bsz, 1))
        eps = denoise(sd, pre, hp, x, c, ts)
        x = x - beta_infer[i] / torch.sqrt(1 - alpha_infer[i] ** 2.0) * eps
        x = x / torch.sqrt(1 - beta_infer[i])
        if i > 0:
            x = x + sigma_infer[i] * noise.pop(0)
    return x


def length_regulator_padded(x, durations, max_length, multiple):
    """LengthRegulator(pad_to_multiple_of=multiple).forward (model.py:349-370)"""
    reps = [torch.repeat_interleave(x[i], durations[i].long(), dim=0) for i in range(x.shape[0])]
    lengths = torch.tensor([r.shape[0] for r in reps]).long()
    ml = min(int(lengths.max()), int(max_length))
    ml = int(math.ceil(ml / multiple) * multiple)
    mask = ~(torch.arange(ml).expand(len(lengths), ml) < lengths.unsqueeze(1))
    out = torch.zeros(x.shape[0], ml, x.shape[2], dtype=x.dtype)
    for i, r in enumerate(reps):
        n = min(r.shape[0], ml)
        out[i, :n] = r[:n]
    return out, mask


def adaptor(sd, cfg, x, src_mask, targets, inference, noise, steps=None, jitter=None, n_steps=4, duration_rounded=None):
    """FastDiffVarianceAdaptor.forward (:83-138).  cfg: variances, variance_nlayers, duration_nlayers, depthwise flags,
    stats, max_length.  noise: list consumed in draw order."""
    noise = list(noise)
    hyper = hyperparams()
    dp = {"nlayers": cfg["duration_nlayers"], "depthwise": cfg["duration_depthwise_conv"]}
    res = {}
    if not inference:
        duration = targets["duration"] + 1 + jitter * 0.49
        duration = (torch.log(duration) - 1.08) / 0.96
        z = noise.pop(0)
        xt = q_sample(duration, z, steps["duration"], hyper)
        dpred = denoise(sd, "duration_predictor.", dp, xt, x, steps["duration"].float()[:, None], mask=src_mask)
        dz = z
        dur = targets["duration"]
    else:
        dpred = predictor_inference(sd, "duration_predictor.", dp, x, n_steps, noise, hyper)
        dz = None
        dpred = dpred * 0.96 + 1.08
        if duration_rounded is None:
            dur = torch.clamp(torch.round(torch.exp(dpred) - 1), min=0).int()
            for i in range(len(dur)):
                if dur[i][~src_mask[i]].sum() <= (~src_mask[i]).sum() // 2:
                    dur[i][~src_mask[i]] = 1
                dur[i][src_mask[i]] = 0
        else:
            dur = duration_rounded
    x, tgt_mask = length_regulator_padded(x, dur, cfg["max_length"], 64)
    out_val = None
    for i, var in enumerate(cfg["variances"]):
        pre = f"encoders.{var}."
        vp = {"nlayers": cfg["variance_nlayers"][i], "depthwise": cfg["variance_depthwise_conv"]}
        st = cfg["stats"][var]
        if not inference:
            tgt = targets[f"variances_{var}"]
            z = noise.pop(0)
            xt = q_sample(tgt, z, steps[var], hyper)
            pred = denoise(sd, pre + "predictor.", vp, xt, x, steps[var].float()[:, None], mask=tgt_mask)
            emb = sd[pre + "embedding.weight"][torch.bucketize(tgt * st["std"] + st["mean"], sd[pre + "bins"])]
        else:
            pred = predictor_inference(sd, pre + "predictor.", vp, x, n_steps, noise, hyper)
            z = None
            emb = sd[pre + "embedding.weight"][torch.bucketize(pred * st["std"] + st["mean"], sd[pre + "bins"])]
        res[f"variances_{var}"] = pred
        res[f"variances_{var}_z"] = z
        if out_val is None:
            out_val = emb
        else:
            out_val = out_val + emb
            x = x + emb
    res.update(x=x, duration_prediction=dpred, duration_z=dz, duration_rounded=dur, tgt_mask=tgt_mask, out=out_val)
    return res
