"""CPU checks of the caller-facing boundary: the argparse path of litfass/train.py:30-93 + :220-262
driven against this repo's FastSpeech2, dataset construction / cache (reference fastspeech2.py:167-228), the
dataloaders (:1308-1323), the `litfass.fastspeech2` import shim, and the optimizer-state / gradient-buffer
contracts of the fused AdamW.  No kernel is launched."""
import inspect
import os
import subprocess
import sys
from argparse import ArgumentParser

import pytest
import torch

from lightningfastspeech2_b200 import configs
from lightningfastspeech2_b200.fastspeech2 import boundary
from lightningfastspeech2_b200.fastspeech2.fastspeech2 import FastSpeech2, FusedAdamW

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"


class FakeTTSDataset(torch.utils.data.Dataset):
    """what FastSpeech2 reads from a built TTSDataset (reference fastspeech2.py:236-245, :1308-1323)"""

    speaker_type = "dvector"

    def __init__(self, raw=None, n=7, **kwargs):
        self.raw, self.kwargs, self.n = raw, kwargs, n
        self.stats = {v: {"min": -3.0, "max": 3.0, "mean": 0.0, "std": 1.0} for v in ("pitch", "energy", "snr")}
        self.phone2id = {f"p{i}": i for i in range(80)}
        self.speaker2dvector = {"spk": [0.0] * 256}
        self.sorted = False

    def create_validation_dataset(self, valid_raw, **kwargs):
        return FakeTTSDataset(valid_raw, n=3, **kwargs)

    def sort_by_duration(self):
        self.sorted = True

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return {"phones": torch.full((3 + i,), 1 + i, dtype=torch.int64), "speaker": torch.zeros(256)}

    def _collate_fn(self, data):
        tp = max(len(d["phones"]) for d in data)
        phones = torch.zeros(len(data), tp, dtype=torch.int64)
        for j, d in enumerate(data):
            phones[j, : len(d["phones"])] = d["phones"]
        return {"phones": phones, "speaker": torch.stack([d["speaker"] for d in data])}


class Raw:
    def __init__(self, h):
        self.hash = h


def _train_py_parser():
    """the parser litfass/train.py:30-93 builds around the two FastSpeech2 hooks (Trainer / wandb flags left out)"""
    parser = ArgumentParser()
    parser.add_argument("--dataset_cache_path", type=str, default="../dataset_cache")
    parser.add_argument("--no_cache", type=boundary.str2bool, default=False)
    parser.add_argument("--train_target_path", type=str, nargs="+", default=["../data/train-clean-360-aligned"])
    parser.add_argument("--valid_target_path", type=str, default="../data/dev-clean-aligned")
    parser = FastSpeech2.add_model_specific_args(parser)
    parser = FastSpeech2.add_dataset_specific_args(parser)
    parser.add_argument("--from_checkpoint", type=str, default=None)
    return parser


def test_argparse_path_of_train_py_constructs_the_model():
    parser = _train_py_parser()
    # flags of the reference's scripts/train.sh that concern the path (its FastDiff options are out of scope)
    args = parser.parse_args(
        "--batch_size 4 --layer_dropout 0.00 --duration_dropout 0.1 --variance_dropout 0.1 0.1 0.1 0.1 "
        "--encoder_hidden 256 --encoder_conv_filter_size 1024 --variance_filter_size 256 --duration_filter_size 256 "
        "--decoder_hidden 256 --decoder_conv_filter_size 1024 --encoder_head 2 --decoder_head 2 "
        "--variance_loss_weights 1 1 1 1 --duration_loss_weight 1 --duration_nlayers 5 "
        "--variances pitch energy snr --variance_levels frame frame frame --variance_transforms none none none "
        "--variance_losses mse mse mse --variance_early_stopping none --decoder_layers 6 "
        "--decoder_kernel_sizes 9 9 9 9 9 9 --speaker_embedding_every_layer False "
        "--prior_embedding_every_layer False --speaker_type dvector --train_min_samples_per_speaker 50 "
        "--sort_data_by_length True --train_pad_to_multiple_of 64 --fastdiff_variances False --num_workers 0".split())
    var_args = vars(args)
    # train.py:109-119 and :220-222
    train_ds_kwargs = {k.replace("train_", ""): v for k, v in var_args.items() if k.startswith("train_")}
    valid_ds_kwargs = {k.replace("valid_", ""): v for k, v in var_args.items() if k.startswith("valid_")}
    assert train_ds_kwargs["min_samples_per_speaker"] == 50 and train_ds_kwargs["pad_to_multiple_of"] == 64
    assert {"max_entries", "stat_entries", "fmin", "fmax", "pitch_quality", "source_phoneset", "shuffle_seed",
            "overwrite_stats", "overwrite_stats_if_missing"} <= set(train_ds_kwargs)
    assert {"max_entries", "shuffle_seed", "nexamples", "example_directory"} <= set(valid_ds_kwargs)
    model_args = {k: v for k, v in var_args.items() if k in inspect.signature(FastSpeech2).parameters}
    assert "lr" in model_args and model_args["lr"] == 2e-4 and "dataset_cache_path" not in model_args
    model = FastSpeech2(FakeTTSDataset(), FakeTTSDataset(n=3), fastdiff_model=None, **model_args)
    hp = model.hparams
    assert hp.decoder_layers == 6 and hp.decoder_kernel_sizes == [9] * 6 and hp.duration_nlayers == 5
    assert len(model.decoder.layers) == 6 and model.batch_size == 4 and hp.sort_data_by_length is True
    assert len(model.phone_embedding.weight) == 80 and model.speaker2dvector == {"spk": [0.0] * 256}
    # Trainer.fit -> train_dataloader / val_dataloader (:1308-1323)
    dl = model.train_dataloader()
    assert model.train_ds.sorted
    batches = list(dl)
    assert len(batches) == 2 and batches[0]["phones"].shape == (4, 6) and batches[0]["phones"].dtype == torch.int64
    assert sum(len(b["phones"]) for b in model.val_dataloader()) == 3


def test_argparse_defaults_are_the_references():
    args = _train_py_parser().parse_args([])
    assert args.lr == 2e-4 and args.variance_loss_weights == [1, 0.1, 0.1, 0.1] and args.fastdiff_variances is False
    assert args.variance_levels == ["frame"] * 4 and args.encoder_kernel_sizes == [5, 25, 13, 9]
    assert args.train_stat_entries == 10_000 and args.valid_shuffle_seed == 42 and args.max_length == 32
    if os.path.isdir(REFERENCE):  # every flag of the reference's own parser exists here with the same default
        src = open(os.path.join(REFERENCE, "litfass/fastspeech2/fastspeech2.py")).read()
        import re

        flags = re.findall(r'"--(\w+)"', src[src.index("def add_model_specific_args"):src.index("def add_dataset_specific_args")])
        assert flags and set(flags) == {n for n, _ in boundary._MODEL_ARGS}


def test_raw_datasets_are_wrapped_and_cached(tmp_path, monkeypatch):
    monkeypatch.setattr(boundary, "_tts_dataset_class", lambda: FakeTTSDataset)
    kw = dict(configs.C2, num_workers=0)
    model = FastSpeech2([Raw("a"), Raw("b")], Raw("v"), train_ds_kwargs={"fmax": 8000}, valid_ds_kwargs={"max_entries": 5},
                        cache_path=str(tmp_path), **kw)
    # the model's feature settings are forced into the dataset kwargs (:169-181)
    assert model.train_ds.kwargs["variances"] == ["pitch", "energy"] and model.train_ds.kwargs["hop_length"] == 256
    assert model.train_ds.kwargs["fmax"] == 8000 and model.valid_ds.kwargs == {"max_entries": 5}
    files = sorted(p.name for p in tmp_path.iterdir())
    assert len(files) == 2 and files[0].startswith("train-full-") and files[1].startswith("valid-full-")
    # second construction: served from the pickle cache (TTSDataset is not called again)
    monkeypatch.setattr(boundary, "_tts_dataset_class", lambda: (lambda *a, **k: pytest.fail("cache miss")))
    again = FastSpeech2([Raw("a"), Raw("b")], Raw("v"), train_ds_kwargs={"fmax": 8000}, valid_ds_kwargs={"max_entries": 5},
                        cache_path=str(tmp_path), **kw)
    assert again.train_ds.hash == model.train_ds.hash and again.stats == model.stats


def test_raw_dataset_without_the_reference_package_raises(monkeypatch):
    monkeypatch.setattr(boundary, "_tts_dataset_class", lambda: None)
    with pytest.raises(ImportError, match="TTSDataset"):
        FastSpeech2(Raw("a"), **dict(configs.C2, num_workers=0))


def test_litfass_import_shim_resolves_to_this_repo():
    shim = os.path.join(ROOT, "lightningfastspeech2_b200", "shim")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([shim, ROOT] + ([REFERENCE] if os.path.isdir(REFERENCE) else [])))
    code = (
        "import litfass, sys\n"
        "from litfass.fastspeech2.fastspeech2 import FastSpeech2, NoamLR, FastSpeech2Loss\n"
        "from litfass.fastspeech2.model import (ConformerEncoderLayer, PositionalEncoding, VarianceAdaptor, PriorEmbedding,\n"
        "    SpeakerEmbedding, VarianceConvolutionLayer, VariancePredictor, VarianceEncoder, LengthRegulator, Transpose)\n"
        "from litfass.fastspeech2.loss import FastSpeech2Loss as L2\n"
        "from litfass.fastspeech2.noam import NoamLR as N2\n"
        "assert FastSpeech2.__module__ == 'lightningfastspeech2_b200.fastspeech2.fastspeech2', FastSpeech2.__module__\n"
        "assert VarianceAdaptor.__module__ == 'lightningfastspeech2_b200.fastspeech2.model'\n"
        "assert L2 is FastSpeech2Loss and N2 is NoamLR\n"
        "assert hasattr(FastSpeech2, 'add_model_specific_args') and hasattr(FastSpeech2, 'load_from_checkpoint')\n"
        "import os\n"
        "if os.path.isdir('/root/reference'):\n"
        "    from litfass.third_party.argutils import str2bool\n"  # other sub-packages still come from the reference
        "    assert '/root/reference' in sys.modules['litfass.third_party.argutils'].__file__\n"
        "    assert str2bool('yes') is True\n"
        "print('ok')\n")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def _small_cpu_model():
    kw = configs.PRESETS["TINY_DW"]
    stats = {v: {"min": -3.0, "max": 3.0, "mean": 0.0, "std": 1.0} for v in kw["variances"]}
    return FastSpeech2(stats=stats, phone2id={f"p{i}": i for i in range(80)}, num_workers=0, **kw)


def test_fused_adamw_state_dict_is_torch_adamw_compatible():
    torch.manual_seed(0)
    model = _small_cpu_model()
    opt = FusedAdamW(model, lr=1e-4)
    assert opt.state_dict()["state"] == {}  # nothing stepped yet
    opt.exp_avg.normal_()
    opt.exp_avg_sq.uniform_()
    opt.step_count = 7
    sd = opt.state_dict()
    params = [p for p in model.parameters() if p.requires_grad]
    assert len(sd["state"]) == len(params) and sd["param_groups"][0]["params"] == list(range(len(params)))
    # -> torch's AdamW (the reference's optimizer, fastspeech2.py:1166-1173) accepts it
    ref = torch.optim.AdamW(params, lr=1e-4, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.01)
    ref.load_state_dict({"state": sd["state"], "param_groups": [dict(ref.state_dict()["param_groups"][0])]})
    st = ref.state[params[3]]
    assert float(st["step"]) == 7 and torch.equal(st["exp_avg"], sd["state"][3]["exp_avg"])
    # <- and a torch AdamW state resumes here: moments and the step count land in the flat buffers
    opt2 = FusedAdamW(model, lr=1e-4)
    opt2.load_state_dict(ref.state_dict())
    assert opt2.step_count == 7
    for o, n, _ in opt._slots():  # (the alignment gaps between parameters are not state)
        assert torch.equal(opt2.exp_avg[o:o + n], opt.exp_avg[o:o + n])
        assert torch.equal(opt2.exp_avg_sq[o:o + n], opt.exp_avg_sq[o:o + n])
    with pytest.raises(ValueError):
        bad = ref.state_dict()
        bad["state"].pop(0)
        opt2.load_state_dict(bad)


def test_gradients_stay_views_of_the_flat_buffer():
    model = _small_cpu_model()
    flat_p, flat_g = model.flatten_parameters()
    params = [p for p in model.parameters() if p.requires_grad]
    flat_g.fill_(1.0)
    model.zero_grad()  # torch / Lightning default set_to_none=True: here it zeroes in place and keeps the views
    assert float(flat_g.abs().sum()) == 0.0 and all(p.grad is not None for p in params)
    p = params[5]
    p.grad = None                      # someone dropped the view ...
    stray = torch.full_like(p, 2.0)
    p.grad = stray                     # ... and a kernel accumulated into a fresh tensor
    model.rehome_gradients()
    assert p.grad.data_ptr() != stray.data_ptr() and torch.equal(p.grad, stray)
    off = (p.grad.data_ptr() - flat_g.data_ptr()) // 4
    assert torch.equal(flat_g[off:off + p.numel()].view(p.shape), stray)
    super(FastSpeech2, model).zero_grad(set_to_none=True)
    model.rehome_gradients()
    assert all(q.grad is not None and q.grad.data_ptr() >= flat_g.data_ptr() for q in params)


def test_length_bucket_plan_covers_every_utterance_with_its_halo():
    """host logic of model.train_length_buckets (training.plan_length_buckets): every utterance lands in exactly one
    bucket, a bucket keeps its longest utterance + the conv halo but never more than the full batch's tensor, buckets
    come longest first, the LengthRegulator cap of a bucket is its own longest utterance (capped by max_length)"""
    import numpy as np

    from lightningfastspeech2_b200.fastspeech2.training import plan_length_buckets

    rng = np.random.default_rng(3)
    for bsz, ngroups in ((1, 3), (7, 2), (7, 3), (64, 4), (5, 9)):
        nphones = rng.integers(3, 120, size=bsz).tolist()
        nframes = [int(p * rng.integers(1, 9)) for p in nphones]
        tp, cap, h_enc, h_dec = max(nphones), 300, 26, 28
        plan = plan_length_buckets(nphones, nframes, ngroups, tp, cap, h_enc, h_dec)
        assert sorted(i for idx, _, _ in plan for i in idx) == list(range(bsz))
        assert len(plan) <= max(1, min(ngroups, bsz))
        l_full = min(max(nframes), cap)
        longest = [max(nframes[i] for i in idx) for idx, _, _ in plan]
        assert longest == sorted(longest, reverse=True)
        for idx, tp_g, (l, cap_g) in plan:
            assert tp_g == min(tp, max(nphones[i] for i in idx) + h_enc)
            assert cap_g == min(max(nframes[i] for i in idx), cap)
            assert cap_g <= l <= l_full and l == min(cap_g + h_dec, l_full)
        # the bucket holding the batch's longest utterance ends exactly where the full tensor ends
        assert plan[0][2][0] == l_full


def test_train_weight_cache_lifecycle():
    """training.WeightCache (host logic of the train step's per-step weight pack): entries are built once per weight state
    and shared; accumulators are created once per backward and finished exactly once by flush(); a parameter change (its
    _version, or ops.WEIGHTS_EPOCH for writes through raw pointers) drops everything; a forward start drops accumulators an
    aborted backward left behind"""
    import torch

    from lightningfastspeech2_b200 import ops
    from lightningfastspeech2_b200.fastspeech2.training import WeightCache

    model = torch.nn.Linear(4, 3)
    model.compute_mode = "fp32"
    cache = WeightCache()
    cache.validate(model)
    built = []
    w = model.weight

    def make():
        built.append(1)
        return w.detach().clone()

    a = cache.get("planes", w, make)
    b = cache.get("planes", w.view(3, 4), make)          # a view of the parameter: same storage, same shape -> same entry
    assert a is b and len(built) == 1
    assert cache.get("planesT", w, make) is not a and len(built) == 2   # another tag is another entry
    finished = []
    acc1 = cache.accumulator("d_fold", w, lambda: torch.zeros(3), lambda acc: finished.append(acc))
    acc2 = cache.accumulator("d_fold", w, lambda: torch.ones(3), lambda acc: finished.append(acc))
    assert acc1 is acc2                                   # the second bucket adds into the first one's accumulator
    cache.flush()
    assert len(finished) == 1 and finished[0] is acc1
    cache.flush()
    assert len(finished) == 1                             # nothing pending any more
    cache.validate(model)
    assert cache.get("planes", w, make) is a and len(built) == 2        # same weights: still cached
    with torch.no_grad():
        w.add_(1.0)                                       # torch-side update: _version moves
    cache.validate(model)
    assert cache.get("planes", w, make) is not a and len(built) == 3
    epoch = ops.WEIGHTS_EPOCH
    try:
        ops.WEIGHTS_EPOCH += 1                            # what the fused optimizer does after writing through raw pointers
        cache.validate(model)
        cache.get("planes", w, make)
        assert len(built) == 4
    finally:
        ops.WEIGHTS_EPOCH = epoch
    cache.validate(model)
    cache.accumulator("d_fold", w, lambda: torch.zeros(3), lambda acc: finished.append(acc))   # a backward that raised ...
    cache.validate(model)                                 # ... must not leak into the next step
    cache.flush()
    assert len(finished) == 1
