"""FastSpeech2Loss default branches (reference litfass/fastspeech2/loss.py:57-81, 83-213):
masked MSE per variance, masked L1 on mel, masked MSE on log-duration, weighted total.

Soft-DTW, CWT and FastDiff/speaker losses are optional branches off the default path and
raise NotImplementedError."""
import torch
from torch import nn


class FastSpeech2Loss(nn.Module):
    def __init__(self, variances=("energy", "pitch", "snr"), variance_levels=("phone", "phone", "phone"),
                 variance_transforms=("cwt", "none", "none"), variance_losses=("mse", "mse", "mse"), mel_loss="l1",
                 duration_loss="mse", duration_stochastic=False, max_length=4096, loss_alphas=None,
                 soft_dtw_gamma=0.01, soft_dtw_chunk_size=256, fastdiff_loss=None, fastdiff_variances=False):
        super().__init__()
        for name in list(variance_losses) + [mel_loss, duration_loss]:
            if name not in ("mse", "l1"):
                raise NotImplementedError(f"loss '{name}'")
        if duration_stochastic or fastdiff_loss is not None or fastdiff_variances:
            raise NotImplementedError("stochastic-duration / FastDiff losses")
        self.variances = list(variances)
        self.variance_levels = list(variance_levels)
        self.variance_transforms = list(variance_transforms)
        self.variance_losses = list(variance_losses)
        self.mel_loss = mel_loss
        self.duration_loss = duration_loss
        self.max_length = max_length
        self.loss_alphas = dict(loss_alphas or {})

    @staticmethod
    def _masked(pred, truth, kind, mask):
        diff = pred[mask] - truth[mask]
        return diff.abs().mean() if kind == "l1" else (diff * diff).mean()

    def forward(self, result, target, frozen_components=()):
        dev, dt = result["mel"].device, result["mel"].dtype
        valid_src = ~result["src_mask"]
        valid_tgt = ~result["tgt_mask"]
        assert target["mel"].shape[1] <= self.max_length
        losses = {}
        for var, level, transform, kind in zip(self.variances, self.variance_levels, self.variance_transforms,
                                               self.variance_losses):
            if transform == "cwt":
                raise NotImplementedError("cwt loss")
            truth = target[f"variances_{var}"].to(dev, dt)
            if level == "frame":
                truth = truth[:, : int(self.max_length)]
                mask = valid_tgt
            elif level == "phone":
                mask = valid_src
            else:
                raise ValueError(f"Unknown variance level: {level}")
            losses[var] = self._masked(result[f"variances_{var}"], truth, kind, mask)
        m = valid_tgt.unsqueeze(-1).expand_as(result["mel"])
        losses["mel"] = self._masked(result["mel"], target["mel"].to(dev, dt), self.mel_loss, m)
        losses["duration"] = self._masked(result["duration_prediction"],
                                          torch.log(target["duration"].to(dev) + 1).to(dt), self.duration_loss,
                                          valid_src)
        losses["total"] = sum(v * self.loss_alphas[k] for k, v in losses.items()
                              if not any(f in k for f in frozen_components))
        return losses
