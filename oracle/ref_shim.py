"""TEST INFRASTRUCTURE ONLY -- import shim that makes the *unmodified* reference importable.

Only usable where /root/reference exists (the authoring container).  It never travels
to the GPU box; what travels are the golden vectors this shim helps generate
(tests/golden/, made by oracle/make_goldens.py) and the restatement in
oracle/fs2_oracle.py that those vectors pin.

What is stubbed and why (all stubs are for packages that are
absent from this image or for HEAD bugs, none touches the arithmetic of the path):
  * litfass.dataset.cwt           -- imports scipy.signal.cwt (removed from SciPy)
  * pytorch_lightning             -- not installed; LightningModule -> nn.Module
  * litfass.dataset.datasets      -- needs pyworld/librosa/...; FakeTTSDataset carries stats
  * pysdtw                        -- not installed; only used by the soft-DTW loss option
  * litfass.third_party.hifigan   -- vocoder; ctor torch.load()s a CUDA checkpoint
  * nn.TransformerEncoder.forward -- torch>=2 passes is_causal / probes linear1, which
                                     ConformerEncoderLayer (reference model.py:67-116)
                                     deletes; replaced by the torch-1.10 per-layer loop
"""
import argparse
import inspect
import os
import sys
import types

import torch
from torch import nn

REFERENCE_ROOT = os.environ.get("LFS2_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "litfass"))


class _LightningModule(nn.Module):
    """Minimal stand-in for pl.LightningModule (only what fastspeech2.py touches)."""

    def __init__(self):
        super().__init__()
        self.current_epoch = 0
        self._logged = {}

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")

    def save_hyperparameters(self, ignore=()):
        frame = inspect.currentframe().f_back
        args, _, _, values = inspect.getargvalues(frame)
        hp = {k: values[k] for k in args if k not in ("self",) and k not in ignore}
        self._hparams = argparse.Namespace(**hp)

    @property
    def hparams(self):
        return self._hparams

    def log_dict(self, d, **kw):
        self._logged.update(d)


class FakeTTSDataset:
    """Carries exactly what FastSpeech2.__init__ reads from a TTSDataset
    (reference fastspeech2.py:236-245)."""

    def __init__(self, ds=None, **kw):
        self.kw = kw
        variances = kw.get("variances", ["pitch", "energy", "snr"])
        self.stats = {v: {"min": -3.0, "max": 3.0, "mean": 0.0, "std": 1.0} for v in variances}
        nphones = getattr(ds, "nphones", 80)
        self.phone2id = {f"p{i}": i for i in range(nphones)}
        self.speaker_type = kw.get("speaker_type", "dvector")
        self.speaker2dvector = {}


def _plain_encoder_forward(self, src, mask=None, src_key_padding_mask=None, **kw):
    out = src
    for mod in self.layers:
        out = mod(out, src_mask=mask, src_key_padding_mask=src_key_padding_mask)
    if self.norm is not None:
        out = self.norm(out)
    return out


_installed = False


def install():
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    cwt = types.ModuleType("litfass.dataset.cwt")

    class CWT:  # only used when variance_transforms == "cwt" (out of scope)
        pass

    cwt.CWT = CWT
    sys.modules["litfass.dataset.cwt"] = cwt

    try:
        import pytorch_lightning  # noqa: F401
    except Exception:
        pl = types.ModuleType("pytorch_lightning")
        pl.LightningModule = _LightningModule
        sys.modules["pytorch_lightning"] = pl

    ds = types.ModuleType("litfass.dataset.datasets")
    ds.TTSDataset = FakeTTSDataset
    sys.modules["litfass.dataset.datasets"] = ds

    sd = types.ModuleType("pysdtw")

    class SoftDTW(nn.Module):
        def __init__(self, *a, **k):
            super().__init__()

    sd.SoftDTW = SoftDTW
    sys.modules["pysdtw"] = sd

    hf = types.ModuleType("litfass.third_party.hifigan")
    hf.Synthesiser = lambda device=None, model=None: None
    sys.modules["litfass.third_party.hifigan"] = hf

    nn.TransformerEncoder.forward = _plain_encoder_forward
    _installed = True


def import_model():
    install()
    import litfass.fastspeech2.model as m

    return m


def import_fastspeech2():
    install()
    import litfass.fastspeech2.fastspeech2 as f

    return f


def build_reference(kwargs, nphones=80, attach_fastdiff_head=True):
    """Construct the reference FastSpeech2 (CPU) with the given ctor kwargs.
    HEAD quirk 1: forward() calls self.fastdiff_linear unconditionally, so
    a head of the reference's own shape (fastspeech2.py:393-402) is attached."""
    f = import_fastspeech2()

    class _DS:
        pass

    _DS.nphones = nphones
    kw = dict(kwargs)
    kw.setdefault("fastdiff_variances", False)
    kw.setdefault("speaker_type", "dvector")
    model = f.FastSpeech2(train_ds=_DS(), **kw)
    if attach_fastdiff_head:
        d = model.hparams.decoder_hidden
        model.fastdiff_linear = nn.Sequential(nn.Linear(d, d), nn.Linear(d, model.hparams.n_mels))
    return model
