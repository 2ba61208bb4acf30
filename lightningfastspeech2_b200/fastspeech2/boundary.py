"""Caller-facing pieces of ``litfass.fastspeech2.fastspeech2.FastSpeech2`` that sit around the hot path:
the argparse surface ``litfass/train.py:73-74`` calls (reference fastspeech2.py:1184-1306), the dataset
construction of the constructor (:167-228, including the pickle cache) and the two dataloaders (:1308-1323).

None of this computes anything; it exists so that ``train.py`` / ``generate.py`` drive this repo's module exactly
like the reference's.  ``TTSDataset`` itself (feature extraction, alignments: litfass/dataset/datasets.py) is out
of scope and is imported lazily from an installed ``litfass`` when raw alignment datasets are passed in.
"""
import argparse
import hashlib
import json
import multiprocessing
import pickle
from copy import copy
from pathlib import Path

num_cpus = multiprocessing.cpu_count()


def str2bool(v):
    """reference litfass/third_party/argutils/__init__.py:3-11"""
    if isinstance(v, bool):
        return v
    if v.lower() in ("yes", "true", "t", "y", "1"):
        return True
    if v.lower() in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("Boolean value expected.")


# (flag, kwargs) in the reference's order, names, types and DEFAULTS (fastspeech2.py:1186-1289).  Note that several
# argparse defaults differ from the constructor's (lr 2e-4 vs 1e-4, variance_loss_weights, fastdiff_variances False):
# both are kept as the reference has them.
_MODEL_ARGS = [
    ("lr", dict(type=float, default=2e-04)),
    ("warmup_steps", dict(type=int, default=4000)),
    ("batch_size", dict(type=int, default=6)),
    ("speaker_type", dict(type=str, default="dvector")),
    ("min_length", dict(type=float, default=0.5)),
    ("max_length", dict(type=float, default=32)),
    ("augment_duration", dict(type=float, default=0.1)),
    ("layer_dropout", dict(type=float, default=0.1)),
    ("variances", dict(nargs="+", type=str, default=["pitch", "energy", "snr"])),
    ("variance_levels", dict(nargs="+", type=str, default=["frame", "frame", "frame", "frame"])),
    ("variance_transforms", dict(nargs="+", type=str, default=["cwt", "none", "none", "none"])),
    ("variance_losses", dict(nargs="+", type=str, default=["mse", "mse", "mse", "mse"])),
    ("variance_nlayers", dict(nargs="+", type=int, default=[5, 5, 5, 5])),
    ("variance_loss_weights", dict(nargs="+", type=float, default=[1, 1e-1, 1e-1, 1e-1])),
    ("variance_kernel_size", dict(nargs="+", type=int, default=[3, 3, 3, 3])),
    ("variance_dropout", dict(nargs="+", type=float, default=[0.5, 0.5, 0.5, 0.5])),
    ("variance_filter_size", dict(type=int, default=256)),
    ("variance_nbins", dict(type=int, default=256)),
    ("variance_depthwise_conv", dict(type=str2bool, default=True)),
    ("duration_nlayers", dict(type=int, default=2)),
    ("duration_loss_weight", dict(type=float, default=5e-1)),
    ("duration_stochastic", dict(type=str2bool, default=False)),
    ("duration_kernel_size", dict(type=int, default=3)),
    ("duration_dropout", dict(type=float, default=0.5)),
    ("duration_filter_size", dict(type=int, default=256)),
    ("duration_depthwise_conv", dict(type=str2bool, default=True)),
    ("duration_loss", dict(type=str, default="mse")),
    ("mel_loss", dict(type=str, default="l1")),
    ("soft_dtw_gamma", dict(type=float, default=0.1)),
    ("soft_dtw_chunk_size", dict(type=int, default=256)),
    ("priors", dict(nargs="+", type=str, default=[])),
    ("mel_loss_weight", dict(type=float, default=1)),
    ("n_mels", dict(type=int, default=80)),
    ("sampling_rate", dict(type=int, default=22050)),
    ("n_fft", dict(type=int, default=1024)),
    ("win_length", dict(type=int, default=1024)),
    ("hop_length", dict(type=int, default=256)),
    ("encoder_hidden", dict(type=int, default=256)),
    ("encoder_head", dict(type=int, default=2)),
    ("encoder_layers", dict(type=int, default=4)),
    ("encoder_dropout", dict(type=float, default=0.1)),
    ("encoder_kernel_sizes", dict(nargs="+", type=int, default=[5, 25, 13, 9])),
    ("encoder_dim_feedforward", dict(type=int, default=None)),
    ("encoder_conformer", dict(type=str2bool, default=True)),
    ("encoder_depthwise_conv", dict(type=str2bool, default=True)),
    ("encoder_conv_filter_size", dict(type=int, default=1024)),
    ("decoder_hidden", dict(type=int, default=256)),
    ("decoder_head", dict(type=int, default=2)),
    ("decoder_layers", dict(type=int, default=4)),
    ("decoder_dropout", dict(type=float, default=0.1)),
    ("decoder_kernel_sizes", dict(nargs="+", type=int, default=[17, 21, 9, 13])),
    ("decoder_dim_feedforward", dict(type=int, default=None)),
    ("decoder_conformer", dict(type=str2bool, default=True)),
    ("decoder_depthwise_conv", dict(type=str2bool, default=True)),
    ("decoder_conv_filter_size", dict(type=int, default=1024)),
    ("valid_nexamples", dict(type=int, default=10)),
    ("valid_example_directory", dict(type=str, default=None)),
    ("variance_early_stopping", dict(type=str, default="none")),
    ("variance_early_stopping_patience", dict(type=int, default=4)),
    ("variance_early_stopping_directory", dict(type=str, default="variance_encoders")),
    ("num_workers", dict(type=int, default=num_cpus)),
    ("speaker_embedding_every_layer", dict(type=str2bool, default=False)),
    ("prior_embedding_every_layer", dict(type=str2bool, default=False)),
    ("priors_gmm", dict(type=str2bool, default=False)),
    ("priors_gmm_max_components", dict(type=int, default=5)),
    ("priors_gmm_min_samples_per_component", dict(type=int, default=20)),
    ("priors_gmm_reg_covar", dict(type=float, default=1e-3)),
    ("priors_gmm_logs", dict(nargs="+", type=int, default=[0, 1, 2, 3])),
    ("dvector_gmm", dict(type=str2bool, default=False)),
    ("fastdiff_schedule", dict(nargs="+", type=int, default=[0.1, 1])),
    ("fastdiff_schedule_start", dict(type=int, default=0)),
    ("fastdiff_schedule_end", dict(type=int, default=30)),
    ("fastdiff_variances", dict(type=str2bool, default=False)),
    ("fastdiff_speakers", dict(type=str2bool, default=False)),
    ("sort_data_by_length", dict(type=str2bool, default=False)),
]

# TTSDataset.add_model_specific_args(parser, split) (reference dataset/datasets.py:1018-1041)
_DATASET_ARGS = [
    ("max_entries", dict(type=int, default=None)),
    ("stat_entries", dict(type=int, default=10_000)),
    ("fmin", dict(type=int, default=0)),
    ("fmax", dict(type=int, default=8000)),
    ("pitch_quality", dict(type=float, default=0.25)),
    ("source_phoneset", dict(type=str, default="arpabet")),
    ("shuffle_seed", dict(type=int, default=42)),
    ("overwrite_stats", dict(type=str2bool, default=False)),
    ("overwrite_stats_if_missing", dict(type=str2bool, default=True)),
    ("min_samples_per_speaker", dict(type=int, default=0)),
    ("pad_to_multiple_of", dict(type=int, default=None)),
]


def add_model_specific_args(parent_parser):
    parser = parent_parser.add_argument_group("FastSpeech2")
    for name, kw in _MODEL_ARGS:
        parser.add_argument(f"--{name}", **kw)
    return parent_parser


def _tts_dataset_class():
    """litfass.dataset.datasets.TTSDataset of an installed reference (feature pipeline: out of scope here), or None"""
    try:
        from litfass.dataset.datasets import TTSDataset  # noqa: PLC0415

        return TTSDataset
    except Exception:  # noqa: BLE001 - the reference's dataset stack needs a dozen audio packages
        return None


def add_dataset_specific_args(parent_parser):
    """reference fastspeech2.py:1299-1306: the train split's TTSDataset flags + two valid-split flags"""
    cls = _tts_dataset_class()
    if cls is not None:
        parent_parser = cls.add_model_specific_args(parent_parser, "train")
    else:
        parser = parent_parser.add_argument_group("train Dataset")
        for name, kw in _DATASET_ARGS:
            parser.add_argument(f"--train_{name}", **kw)
    parser = parent_parser.add_argument_group("Valid Dataset")
    parser.add_argument("--valid_max_entries", type=int, default=None)
    parser.add_argument("--valid_shuffle_seed", type=int, default=42)
    return parent_parser


def _is_built(ds):
    """a finished TTSDataset (or a stand-in for one) exposes what the model reads from it (:236-245)"""
    return hasattr(ds, "stats") and hasattr(ds, "phone2id")


def _cached(path, build):
    if path.exists():
        with path.open("rb") as f:
            return pickle.load(f)
    ds = build()
    with open(path, "wb") as f:
        pickle.dump(ds, f)
    return ds


def build_datasets(train_ds, valid_ds, train_ds_kwargs, valid_ds_kwargs, cache_path, model_kwargs):
    """Constructor part reference fastspeech2.py:167-228: wrap raw alignment datasets in ``TTSDataset`` with the model's
    feature settings, derive the validation dataset from the train one, both behind the md5-keyed pickle cache.
    Already-built datasets (anything exposing ``stats`` and ``phone2id``) pass through untouched.
    -> (train dataset or None, valid dataset or None)"""
    out_train = out_valid = None
    if train_ds is not None:
        if _is_built(train_ds):
            out_train = train_ds
        else:
            cls = _tts_dataset_class()
            if cls is None:
                raise ImportError(
                    "FastSpeech2(train_ds=<raw alignment dataset>) builds a litfass.dataset.datasets.TTSDataset, which "
                    "needs the reference's dataset package on sys.path (it is outside the accelerated path); pass a "
                    "built dataset, or stats= / phone2id=, to construct the model without it")
            kw = dict(train_ds_kwargs or {})
            for k in ("speaker_type", "min_length", "max_length", "augment_duration", "variances", "variance_levels",
                      "variance_transforms", "priors", "n_mels", "sampling_rate", "n_fft", "win_length", "hop_length"):
                kw[k] = model_kwargs[k]
            if cache_path is not None:
                hashes = [x.hash for x in train_ds] if isinstance(train_ds, list) else [train_ds.hash]
                key = copy(kw)
                key.update({"hashes": hashes})
                ds_hash = hashlib.md5(json.dumps(key, sort_keys=True).encode("utf-8")).hexdigest()

                def make():
                    ds = cls(train_ds, **kw)
                    ds.hash = ds_hash
                    return ds

                out_train = _cached(Path(cache_path) / f"train-full-{ds_hash}.pt", make)
            else:
                out_train = cls(train_ds, **kw)
    if valid_ds is not None:
        if _is_built(valid_ds):
            out_valid = valid_ds
        else:
            if out_train is None or not hasattr(out_train, "create_validation_dataset"):
                raise ValueError("valid_ds needs a TTSDataset train_ds to derive the validation dataset from (:205-228)")
            kw = dict(valid_ds_kwargs or {})
            if cache_path is not None:
                key = copy(kw)
                key.update({"hashes": [out_train.hash, valid_ds.hash]})
                ds_hash = hashlib.md5(json.dumps(key, sort_keys=True).encode("utf-8")).hexdigest()

                def make_valid():
                    ds = out_train.create_validation_dataset(valid_ds, **kw)
                    ds.hash = ds_hash
                    return ds

                out_valid = _cached(Path(cache_path) / f"valid-full-{ds_hash}.pt", make_valid)
            else:
                out_valid = out_train.create_validation_dataset(valid_ds, **kw)
    return out_train, out_valid


def dataloader(ds, batch_size, num_workers):
    """reference :1311-1323: DataLoader(ds, batch_size, collate_fn=ds._collate_fn, num_workers)"""
    from torch.utils.data import DataLoader  # noqa: PLC0415

    return DataLoader(ds, batch_size=batch_size, collate_fn=ds._collate_fn, num_workers=num_workers)
