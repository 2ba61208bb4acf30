"""Host-side mirror of ``litfass.fastspeech2.fastdiff_variances`` (reference litfass/fastspeech2/fastdiff_variances.py:8-320):
the FastDiff variance adaptor -- duration and frame-level variances predicted by small DDPMs whose noise-predicting
networks are the SAME ``VarianceConvolutionLayer`` stacks as the plain adaptor's predictors.

Same class names, constructor signatures and state_dict keys as the reference (``predictor.linear_in``, ``fc_t1``,
``fc_t2``, ``linear_noise``, ``layers.<l>.layers...``, ``linear``, ``bins``, ``embedding``).  The stacks run on the
predictor kernels of liblfs2.so (depthwise conv + tcgen05 GEMM with ReLU + LayerNorm epilogue + row-dot head); the front end
(scalar track lifted to d channels + condition + diffusion-step embedding) and the DDPM updates are the streaming
kernels of csrc/diffusion.cu.  Host code only does what the reference also does on the host: the noise-schedule arithmetic
(a handful of scalars) and drawing the Gaussian noise (``noise=`` injects recorded draws for parity runs).

Implemented: inference (reverse diffusion, N in {3, 4, 6, 8, 200, 1000}) and the teacher-forced forward (noise prediction at
a random step; forward values only -- training this adaptor needs backward kernels that are not written, so the train step
raises).  ``FastDiffSpeakerGenerator`` (d-vector diffusion) is not part of the mel path and is not mirrored.
"""
import numpy as np
import torch
from torch import nn

from .. import ops
from .model import LengthRegulator, VarianceConvolutionLayer

# reverse-step noise schedules of FastDiffVariancePredictor.inference (fastdiff_variances.py:240-262)
INFERENCE_SCHEDULES = {
    8: [6.689325005027058e-07, 1.0033881153503899e-05, 0.00015496854030061513, 0.002387222135439515,
        0.035597629845142365, 0.3681158423423767, 0.4735414385795593, 0.5],
    6: [1.7838445955931093e-06, 2.7984189728158526e-05, 0.00043231004383414984, 0.006634317338466644,
        0.09357017278671265, 0.6000000238418579],
    4: [3.2176e-04, 2.5743e-03, 2.5376e-02, 7.0414e-01],
    3: [9.0000e-05, 9.0000e-03, 6.0000e-01],
}


def compute_hyperparams_given_schedule(beta):
    """third_party/fastdiff/module/util.py:276-302 (fp32 CPU tensors; note: ``beta`` is NOT modified, alpha/sigma are new)"""
    t_steps = len(beta)
    alpha = 1 - beta
    sigma = beta + 0
    for t in range(1, t_steps):
        alpha[t] *= alpha[t - 1]
        sigma[t] *= (1 - alpha[t - 1]) / (1 - alpha[t])
    return {"T": t_steps, "beta": beta, "alpha": torch.sqrt(alpha), "sigma": torch.sqrt(sigma)}


def map_noise_scale_to_time_step(alpha_infer, alpha):
    """util.py:305-315: fractional training step whose noise level equals alpha_infer (vectorised search, same result)"""
    if alpha_infer < alpha[-1]:
        return len(alpha) - 1
    if alpha_infer > alpha[0]:
        return 0
    hit = torch.nonzero((alpha[1:] <= alpha_infer) & (alpha_infer <= alpha[:-1]))
    if len(hit) == 0:
        return -1
    t = int(hit[0])
    step_diff = alpha[t] - alpha_infer
    step_diff = step_diff / (alpha[t] - alpha[t + 1])
    return t + step_diff.item()


def inference_plan(diffusion_hyperparams, n_steps, dtype=torch.float32):
    """The scalars of util.py:158-228 (sampling_given_noise_schedule) for a reverse-step count: per kept step n
    (step embedding value, coefficient of eps, 1 / sqrt(1 - beta), sigma)."""
    if n_steps == 1000:
        schedule = torch.linspace(0.000001, 0.01, 1000)
    elif n_steps == 200:
        schedule = torch.linspace(0.0001, 0.02, 200)
    elif n_steps in INFERENCE_SCHEDULES:
        schedule = torch.FloatTensor(INFERENCE_SCHEDULES[n_steps]).to(dtype)
    else:
        raise ValueError("Reverse step should be 3, 4, 6, 8, 200 or 1000.")
    alpha = diffusion_hyperparams["alpha"]
    n = len(schedule)
    beta_infer = schedule
    alpha_infer = 1 - beta_infer
    sigma_infer = beta_infer + 0
    for i in range(1, n):
        alpha_infer[i] *= alpha_infer[i - 1]
        sigma_infer[i] *= (1 - alpha_infer[i - 1]) / (1 - alpha_infer[i])
    alpha_infer = torch.sqrt(alpha_infer)
    sigma_infer = torch.sqrt(sigma_infer)
    steps = []
    for i in range(n):
        step = map_noise_scale_to_time_step(alpha_infer[i], alpha)
        if step >= 0:
            steps.append(step)
    steps = torch.FloatTensor(steps)
    plan = []
    for i in range(len(steps)):
        plan.append((float(steps[i]), float(beta_infer[i] / torch.sqrt(1 - alpha_infer[i] ** 2.0)),
                     float(1.0 / torch.sqrt(1 - beta_infer[i])), float(sigma_infer[i])))
    return plan


def _draw(noise, shape, device):
    """next Gaussian draw: from the injected list (parity runs) or torch's generator on the device"""
    if noise is not None:
        z = noise.pop(0).to(device=device, dtype=torch.float32).contiguous()
        if tuple(z.shape) != tuple(shape):
            raise ValueError(f"injected noise {tuple(z.shape)} does not match {tuple(shape)}")
        return z
    return torch.randn(shape, device=device, dtype=torch.float32)


class FastDiffVariancePredictor(nn.Module):
    """reference fastdiff_variances.py:141-282"""

    def __init__(self, nlayers, in_channels, filter_size, kernel_size, dropout, depthwise, diffusion_hyperparams,
                 diffusion_step_embed_dim_in, diffusion_step_embed_dim_mid, diffusion_step_embed_dim_out):
        super().__init__()
        self.diffusion_hyperparams = diffusion_hyperparams
        self.linear_in = nn.Linear(1, in_channels)
        self.layers = nn.Sequential(*[VarianceConvolutionLayer(in_channels, filter_size, kernel_size, dropout, depthwise)
                                      for _ in range(nlayers)])
        self.diffusion_step_embed_dim_in = diffusion_step_embed_dim_in
        self.fc_t = nn.ModuleList()
        self.fc_t1 = nn.Linear(diffusion_step_embed_dim_in, diffusion_step_embed_dim_mid)
        self.fc_t2 = nn.Linear(diffusion_step_embed_dim_mid, diffusion_step_embed_dim_out)
        self.linear = nn.Linear(filter_size, 1)
        self.linear_noise = nn.Linear(diffusion_step_embed_dim_out, in_channels)

    def _noise_embed(self, ts):
        """ts (B) fp32 -> (B, d): linear_noise(swish(fc_t2(swish(fc_t1(sin/cos embedding)))))   (:192-204)"""
        e = ops.diffusion_step_embed(ts.contiguous(), self.diffusion_step_embed_dim_in)
        e = ops.swish_(ops.linear(e, self.fc_t1.weight, self.fc_t1.bias))
        e = ops.swish_(ops.linear(e, self.fc_t2.weight, self.fc_t2.bias))
        return ops.linear(e, self.linear_noise.weight, self.linear_noise.bias)

    def denoise(self, xt, c, ts, mask=None):
        """epsilon_theta(x_t, c, t): xt (B, L) noisy track, c (B, L, d) condition channels-last, ts (B) -> (B, L)"""
        if self.training and any(layer.layers[3].p > 0 for layer in self.layers):
            raise NotImplementedError("FastDiff predictors: dropout in training mode is not implemented by the CUDA path")
        ne = self._noise_embed(ts)
        z = ops.diffusion_input(xt.contiguous(), self.linear_in.weight.reshape(-1), self.linear_in.bias, c.contiguous(), ne)
        nl = len(self.layers)
        for i, layer in enumerate(self.layers):
            z = layer(z, out="planes" if i + 1 < nl else "f32")
        if isinstance(z, ops.Planes):
            z = ops.merge_planes(z)
        return ops.rowdot_mask(z, self.linear.weight, self.linear.bias, mask)

    def forward(self, x, c, ts=None, mask=None, noise=None, steps=None):
        """Reference signature: x (B, L) clean (ts None) or noisy track, c (B, C, L) channels-FIRST condition.
        ts None: draw a step per utterance and Gaussian noise, return (noise prediction, z) (:177-190, 218-219);
        otherwise return the noise prediction.  ``steps`` (B int64) / ``noise`` inject the draws for parity runs."""
        cl = c.transpose(1, 2).contiguous() if c.dim() == 3 else c.unsqueeze(0).transpose(1, 2).contiguous()
        dev = cl.device
        x = x.to(dev, torch.float32)
        if ts is not None:
            return self.denoise(x, cl, ts.reshape(-1).to(dev, torch.float32), mask)
        bsz = cl.shape[0]
        alpha = self.diffusion_hyperparams["alpha"]
        t_idx = steps if steps is not None else torch.randint(self.diffusion_hyperparams["T"], size=(bsz,))
        t_idx = t_idx.reshape(-1).cpu()
        z = _draw(noise, (x.shape[0], 1, x.shape[1]), dev)                   # the reference draws it as (B, 1, L) (:179-184)
        a = alpha[t_idx].to(dev, torch.float32).contiguous()
        delta = (1 - alpha[t_idx] ** 2.0).sqrt().to(dev, torch.float32).contiguous()
        noisy = ops.diffusion_mix(x.contiguous(), a=a, y=z.view(x.shape), e=delta)        # q(x_t | x_0)
        return self.denoise(noisy, cl, t_idx.to(dev, torch.float32), mask), z

    def inference(self, c, N=4, noise=None):
        """Reverse diffusion conditioned on c (B, L, d) channels-last (the reference transposes it itself, :229) -> (B, L)"""
        c = c.contiguous()
        bsz, length, _ = c.shape
        dev = c.device
        plan = inference_plan(self.diffusion_hyperparams, N)
        x = _draw(noise, (bsz, length), dev)
        ones = torch.ones(bsz, device=dev, dtype=torch.float32)
        for n in range(len(plan) - 1, -1, -1):
            step, k_eps, inv_sqrt, sigma = plan[n]
            eps = self.denoise(x, c, ones * step)
            z = _draw(noise, (bsz, length), dev) if n > 0 else None
            # x <- (x - k eps) / sqrt(1 - beta) [+ sigma z]      (util.py:224-228)
            x = ops.diffusion_mix(x, y=eps, e=ones * (-k_eps), s=ones * inv_sqrt, z=z, g=ones * sigma if n > 0 else None)
        return x


class FastDiffVarianceEncoder(nn.Module):
    """reference fastdiff_variances.py:284-341"""

    def __init__(self, nlayers, in_channels, filter_size, kernel_size, dropout, depthwise, min, max, mean, std, nbins,
                 diffusion_hyperparams, diffusion_step_embed_dim_in, diffusion_step_embed_dim_mid,
                 diffusion_step_embed_dim_out):
        super().__init__()
        self.predictor = FastDiffVariancePredictor(nlayers, in_channels, filter_size, kernel_size, dropout, depthwise,
                                                   diffusion_hyperparams, diffusion_step_embed_dim_in,
                                                   diffusion_step_embed_dim_mid, diffusion_step_embed_dim_out)
        self.bins = nn.Parameter(torch.linspace(min, max, nbins - 1), requires_grad=False)
        self.embedding = nn.Embedding(nbins, in_channels)
        self.mean = mean
        self.std = std

    def forward(self, x, tgt, mask, N=4, control=1.0, noise=None, steps=None):
        """x: (B, C, L) channels-first with a target (training call of the reference), (B, L, C) without (inference).
        -> ((noise prediction, z), embedding)  or  (prediction, embedding); embedding (B, L, C)"""
        if tgt is not None:
            tgt = tgt.to(self.bins.device, torch.float32).contiguous()
            noise_pred, z = self.predictor(tgt, x, mask=mask, noise=noise, steps=steps)
            emb = torch.zeros(tgt.shape + (self.embedding.weight.shape[1],), device=tgt.device, dtype=torch.float32)
            ops.bucket_embed_add_(emb, tgt, self.std, self.mean, self.bins, self.embedding.weight)
            return (noise_pred, z), emb
        prediction = self.predictor.inference(x, N=N, noise=noise)
        emb = torch.zeros(x.shape, device=x.device, dtype=torch.float32)
        ops.bucket_embed_add_(emb, prediction, self.std, self.mean, self.bins, self.embedding.weight)
        if control != 1.0:
            prediction = ops.diffusion_mix(prediction, a=torch.full((prediction.shape[0],), float(control),
                                                                    device=prediction.device))
        return prediction, emb


class FastDiffVarianceAdaptor(nn.Module):
    """reference fastdiff_variances.py:8-138 (1-d, frame-level variances only)"""

    def __init__(self, stats, variances, variance_nlayers, variance_kernel_size, variance_dropout, variance_filter_size,
                 variance_nbins, variance_depthwise_conv, duration_nlayers, duration_kernel_size, duration_dropout,
                 duration_filter_size, duration_depthwise_conv, encoder_hidden, max_length,
                 diffusion_step_embed_dim_in=128, diffusion_step_embed_dim_mid=512, diffusion_step_embed_dim_out=512,
                 beta_0=1e-6, beta_T=0.01, T=1000):
        super().__init__()
        self.max_length = max_length
        self.diffusion_step_embed_dim_in = diffusion_step_embed_dim_in
        self.noise_schedule = torch.linspace(beta_0, beta_T, T)
        self.diffusion_hyperparams = compute_hyperparams_given_schedule(self.noise_schedule)
        dims = (diffusion_step_embed_dim_in, diffusion_step_embed_dim_mid, diffusion_step_embed_dim_out)
        self.duration_predictor = FastDiffVariancePredictor(duration_nlayers, encoder_hidden, duration_filter_size,
                                                            duration_kernel_size, duration_dropout,
                                                            duration_depthwise_conv, self.diffusion_hyperparams, *dims)
        self.length_regulator = LengthRegulator(pad_to_multiple_of=64)
        self.variances = variances
        encoders = {}
        for var in self.variances:
            i = variances.index(var)
            encoders[var] = FastDiffVarianceEncoder(variance_nlayers[i], encoder_hidden, variance_filter_size,
                                                    variance_kernel_size[i], variance_dropout[i], variance_depthwise_conv,
                                                    stats[var]["min"], stats[var]["max"], stats[var]["mean"],
                                                    stats[var]["std"], variance_nbins, self.diffusion_hyperparams, *dims)
        self.encoders = nn.ModuleDict(encoders)

    def forward(self, x, src_mask, targets, inference=False, N=4, noise=None, steps=None, jitter=None, force=None):
        """x (B, Tp, d) encoder output, src_mask (B, Tp) bool.  ``noise`` (list of tensors, consumed in the reference's
        draw order), ``steps`` ({name: (B) int64}) and ``jitter`` ((B, Tp) uniform [0,1) draw of :91) inject the random
        draws; ``force = {"duration_rounded": ...}`` forces the discrete duration decision (parity harness)."""
        steps = steps or {}
        force = force or {}
        x = x.contiguous()
        dev = x.device
        if not inference:
            dur_t = targets["duration"].to(dev)
            u = jitter.to(dev, torch.float32) if jitter is not None else torch.rand(dur_t.shape, device=dev)
            duration = (torch.log(dur_t + 1 + u * 0.49) - 1.08) / 0.96      # (:91-92; a (B, Tp) target transform)
            duration_pred, duration_z = self.duration_predictor(duration.to(torch.float32), x.transpose(1, 2),
                                                                mask=src_mask, noise=noise, steps=steps.get("duration"))
            duration_rounded = dur_t
        else:
            raw = self.duration_predictor.inference(x, N=N, noise=noise)
            duration_z = None
            ones = torch.ones(raw.shape[0], device=dev, dtype=torch.float32)
            duration_pred = ops.diffusion_mix(raw, a=ones * 0.96, add=1.08)                       # (:108)
            if "duration_rounded" in force:
                duration_rounded = force["duration_rounded"].to(dev)
            else:
                # round(exp(p) - 1), clamp >= 0, int32; an utterance whose valid durations sum to <= n_valid // 2 gets all
                # ones; PAD phones get 0 (:109-116) -- the plain adaptor's guard kernel on the PAD-zeroed track
                masked = ops.diffusion_mix(raw, a=ones * 0.96, add=1.08, zero_mask=src_mask)
                duration_rounded = ops.duration_round_guard(masked, src_mask)
        x, tgt_mask = self.length_regulator(x, duration_rounded, self.max_length)
        result = {}
        out_val = None
        for var in self.variances:
            enc = self.encoders[var]
            if not inference:
                tgt = targets[f"variances_{var}"].to(dev, torch.float32)[:, : x.shape[1]].contiguous()
                if tgt.shape[1] < x.shape[1]:  # the collated target is padded to the same multiple of 64 in the reference
                    tgt = torch.nn.functional.pad(tgt, (0, x.shape[1] - tgt.shape[1]))
                (pred, z), out = enc(x.transpose(1, 2), tgt, tgt_mask, noise=noise, steps=steps.get(var))
            else:
                pred, out = enc(x, None, tgt_mask, N=N, noise=noise)
                z = None
            result[f"variances_{var}"] = pred
            result[f"variances_{var}_z"] = z
            if out_val is None:
                out_val = out          # (:125-129: the FIRST variance's embedding is not added to x -- kept as is)
            else:
                out_val = ops.add_(out_val, out)
                x = ops.add_(x, out)
        result["x"] = x
        result["duration_prediction"] = duration_pred
        result["duration_z"] = duration_z
        result["duration_rounded"] = duration_rounded
        result["tgt_mask"] = tgt_mask
        result["out"] = out_val
        return result
