"""FastSpeech2Loss default branches (reference litfass/fastspeech2/loss.py:57-81, 83-213):
masked MSE per variance, masked L1 on mel, masked MSE on log-duration, weighted total.

On CUDA results every loss is one fused kernel (lfs2_masked_loss: value + d(weight*loss)/d pred
in a single pass, no boolean-index gathers, no host sync); the values live in one small device
buffer and ``losses["total"]`` is connected to autograd so ``training_step`` keeps the reference's
contract (return the total, the trainer calls ``.backward()``).

Soft-DTW, CWT and FastDiff/speaker losses are optional branches off the default path and
raise NotImplementedError."""
import torch
from torch import nn

from .. import ops


class _LossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, total, grads, *preds):
        ctx.grads = grads
        return total.clone()

    @staticmethod
    def backward(ctx, gtotal):
        out = []
        for g in ctx.grads:
            out.append(None if g is None else ops.scale_by_(g, gtotal.contiguous()))
        ctx.grads = None
        return (None, None, *out)


class FastSpeech2Loss(nn.Module):
    def __init__(self, variances=("energy", "pitch", "snr"), variance_levels=("phone", "phone", "phone"),
                 variance_transforms=("cwt", "none", "none"), variance_losses=("mse", "mse", "mse"), mel_loss="l1",
                 duration_loss="mse", duration_stochastic=False, max_length=4096, loss_alphas=None,
                 soft_dtw_gamma=0.01, soft_dtw_chunk_size=256, fastdiff_loss=None, fastdiff_variances=False):
        super().__init__()
        for name in list(variance_losses) + [mel_loss, duration_loss]:
            if name not in ("mse", "l1"):
                raise NotImplementedError(f"loss '{name}'")
        if fastdiff_loss is not None or fastdiff_variances:
            raise NotImplementedError("FastDiff losses")
        # the stochastic duration predictor's loss is its own negative log-likelihood (reference loss.py:181-187), i.e. the
        # training direction of the flows, which the CUDA path does not implement: such a model synthesises, but
        # calling the loss raises
        self.duration_stochastic = bool(duration_stochastic)
        self.variances = list(variances)
        self.variance_levels = list(variance_levels)
        self.variance_transforms = list(variance_transforms)
        self.variance_losses = list(variance_losses)
        self.mel_loss = mel_loss
        self.duration_loss = duration_loss
        self.max_length = max_length
        self.loss_alphas = dict(loss_alphas or {})

    def forward(self, result, target, frozen_components=()):
        if self.duration_stochastic:
            raise NotImplementedError("loss of the stochastic duration predictor (training direction of the flows)")
        mel = result["mel"]
        if not mel.is_cuda:
            raise ops._lib.Lfs2Error("FastSpeech2Loss needs CUDA results: there is no CPU path")
        dev = mel.device
        src_mask, tgt_mask = result["src_mask"], result["tgt_mask"]
        assert target["mel"].shape[1] <= self.max_length
        want_grad = torch.is_grad_enabled() and mel.requires_grad
        names = list(self.variances) + ["mel", "duration"]
        buf = torch.zeros(len(names) + 1, device=dev, dtype=torch.float32)  # [losses..., total]
        total = buf[len(names):]
        preds, grads = [], []

        def one(i, name, pred, kind, mask, tgt=None, tgt_i64=None):
            frozen = any(f in name for f in frozen_components)
            w = 0.0 if frozen else float(self.loss_alphas[name])
            g = ops.masked_loss(pred.detach().contiguous(), tgt, mask, kind, w, buf[i:i + 1], None if frozen else total,
                                want_grad=want_grad and not frozen, target_i64=tgt_i64)
            preds.append(pred)
            grads.append(g)

        for i, (var, level, transform, kind) in enumerate(zip(self.variances, self.variance_levels,
                                                              self.variance_transforms, self.variance_losses)):
            if transform == "cwt":
                raise NotImplementedError("cwt loss")
            pred = result[f"variances_{var}"]
            truth = target[f"variances_{var}"].to(dev, torch.float32)
            if level == "frame":
                truth = truth[:, : int(self.max_length)]
                mask = tgt_mask
            elif level == "phone":
                mask = src_mask
            else:
                raise ValueError(f"Unknown variance level: {level}")
            one(i, var, pred, kind, mask, tgt=truth[:, : pred.shape[1]].contiguous())
        nv = len(self.variances)
        # a collated target can be longer than this batch's own longest utterance (sharded batches, padding to a
        # multiple): frames past the prediction are PAD for every utterance here
        one(nv, "mel", mel, self.mel_loss, tgt_mask,
            tgt=target["mel"].to(dev, torch.float32)[:, : mel.shape[1]].contiguous())
        one(nv + 1, "duration", result["duration_prediction"], self.duration_loss, src_mask,
            tgt_i64=target["duration"].to(dev, torch.int64)[:, : result["duration_prediction"].shape[1]].contiguous())
        losses = {name: buf[i] for i, name in enumerate(names)}
        losses["total"] = _LossFn.apply(total, grads, *preds).reshape(()) if want_grad else total.reshape(())
        self.last_buffer = buf  # one D2H copy of this gives every value (training_step logging)
        return losses
