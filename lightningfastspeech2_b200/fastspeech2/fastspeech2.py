"""Host-side mirror of ``litfass.fastspeech2.fastspeech2.FastSpeech2``
(reference litfass/fastspeech2/fastspeech2.py:45-1323): same constructor kwargs, same
``forward(targets, inference=False) -> dict`` contract and result keys (:636-784), same
state_dict layout and checkpoint hooks (:530-634), same optimizer/scheduler recipe
(:1166-1182).  The mel-generation hot path -- Encoder -> VarianceAdaptor -> Decoder -> mel
Linear -- runs entirely in the hand-written sm_100a kernels of liblfs2.so.

Out of scope here: dataset construction (pass a prebuilt dataset object or
``stats=`` / ``phone2id=``), W&B logging, GMM priors, vocoders, FastDiff adaptor, CWT,
stochastic durations.  Those kwargs are accepted and stored in ``hparams`` so callers'
code keeps working, and raise NotImplementedError where they would change the path.
"""
import argparse
import inspect
import multiprocessing

import torch
from torch import nn

from .. import ops
from . import boundary, training
from .loss import FastSpeech2Loss
from .model import (COMPUTE_MODES, ConformerEncoderLayer, PositionalEncoding, PriorEmbedding, SpeakerEmbedding,
                    VarianceAdaptor, _PackCache)
from .noam import NoamLR

try:  # the real Lightning base class when it is installed (it is not in this image)
    from pytorch_lightning import LightningModule as _Base
except Exception:  # pragma: no cover - exercised in this image

    class _Base(nn.Module):
        """Just enough of pl.LightningModule for FastSpeech2 to be driven by a plain loop."""

        def __init__(self):
            super().__init__()
            self.current_epoch = 0
            self.logged = {}

        @property
        def device(self):
            for p in self.parameters():
                return p.device
            return torch.device("cpu")

        def save_hyperparameters(self, ignore=()):
            frame = inspect.currentframe().f_back
            args, _, _, values = inspect.getargvalues(frame)
            self._hparams = argparse.Namespace(**{k: values[k] for k in args if k != "self" and k not in ignore})

        @property
        def hparams(self):
            return self._hparams

        def log_dict(self, d, **kw):
            self.logged.update(d)

        @classmethod
        def load_from_checkpoint(cls, path, strict=True, map_location=None, **kwargs):
            ckpt = torch.load(path, map_location=map_location or "cpu", weights_only=False)
            hp = dict(ckpt.get("hyper_parameters", {}))
            hp.update(kwargs)
            accepted = inspect.signature(cls.__init__).parameters
            model = cls(**{k: v for k, v in hp.items() if k in accepted})
            model.on_load_checkpoint(ckpt)
            model.load_state_dict(ckpt["state_dict"], strict=strict)
            return model


num_cpus = multiprocessing.cpu_count()


class TransformerStack(nn.Module):
    """Stands in for nn.TransformerEncoder(num_layers) as the reference uses it
    (fastspeech2.py:249-286, 347-374): a ``layers`` ModuleList driven in order, norm=None.
    (torch>=2's TransformerEncoder.forward cannot drive ConformerEncoderLayer -- quirk 3.)"""

    def __init__(self, layers):
        super().__init__()
        self.layers = nn.ModuleList(layers)

    def supports_row_limit(self, d):
        return all(mod.supports_row_limit(d) for mod in self.layers)

    def tc_capable(self, d):
        return len(self.layers) > 0 and all(mod.tc_capable(d) for mod in self.layers)

    def wants_f16_plane(self, d):
        """the first block's QKV GEMM runs the 2-pass recipe: its input Planes should carry the fp16 plane"""
        first = self.layers[0] if len(self.layers) else None
        return first is not None and first.compute_mode == "fp32" and "qkv" in first.two_pass_sites and d == 256

    def forward(self, src, mask=None, src_key_padding_mask=None, return_planes=False, row_limit=None):
        """return_planes=True (tensor-core path only) hands the result back as bf16 hi/lo planes,
        the operand format of the next GEMM, instead of materialising an fp32 tensor.
        row_limit = (lengths, extra): rows at or after roundup128(lengths[b] + extra) are neither computed nor
        written by any layer (FastSpeech2.skip_pad_rows)."""
        if mask is None and all(mod.tc_capable(src.shape[-1]) for mod in self.layers):
            xp = src if isinstance(src, ops.Planes) else ops.planes_of(src, want_f16=self.wants_f16_plane(src.shape[-1]))
            if row_limit is not None:
                row_limit = (row_limit[0], row_limit[1], {})  # one tile list per kernel family for the whole stack
            for i, mod in enumerate(self.layers):
                if mod.training and mod.p_drop > 0:
                    raise NotImplementedError("dropout in training mode is not implemented by the CUDA path")
                xp = mod.forward_planes(xp, src_key_padding_mask, row_limit=row_limit, next_f16=i + 1 < len(self.layers))
            return xp if return_planes else ops.merge_planes(xp)
        if row_limit is not None:
            raise NotImplementedError("row limits need the tensor-core path")
        if isinstance(src, ops.Planes):
            raise TypeError("a Planes input needs the tensor-core path")
        out = src
        for mod in self.layers:
            out = mod(out, src_mask=mask, src_key_padding_mask=src_key_padding_mask)
        return ops.planes_of(out) if return_planes else out


class FastSpeech2(_Base):
    def __init__(
        self,
        train_ds=None,
        valid_ds=None,
        lr=1e-04,
        warmup_steps=4000,
        batch_size=6,
        speaker_type="dvector",
        min_length=0.5,
        max_length=32,
        augment_duration=0.1,
        layer_dropout=0.1,
        variances=["pitch", "energy", "snr"],
        variance_levels=["frame", "frame", "frame"],
        variance_transforms=["cwt", "none", "none"],
        variance_losses=["mse", "mse", "mse"],
        variance_nlayers=[5, 5, 5, 5],
        variance_loss_weights=[5e-2, 5e-2, 5e-2, 5e-2],
        variance_kernel_size=[3, 3, 3, 3],
        variance_dropout=[0.5, 0.5, 0.5, 0.5],
        variance_filter_size=256,
        variance_nbins=256,
        variance_depthwise_conv=True,
        duration_nlayers=2,
        duration_loss="mse",
        duration_loss_weight=5e-1,
        duration_stochastic=False,
        duration_kernel_size=3,
        duration_dropout=0.5,
        duration_filter_size=256,
        duration_depthwise_conv=True,
        mel_loss="l1",
        soft_dtw_gamma=0.1,
        soft_dtw_chunk_size=256,
        speaker_embedding_every_layer=False,
        prior_embedding_every_layer=False,
        priors=[],
        mel_loss_weight=1,
        n_mels=80,
        sampling_rate=22050,
        n_fft=1024,
        win_length=1024,
        hop_length=256,
        train_ds_kwargs=None,
        valid_ds_kwargs=None,
        encoder_hidden=256,
        encoder_head=2,
        encoder_layers=4,
        encoder_dropout=0.1,
        encoder_kernel_sizes=[5, 25, 13, 9],
        encoder_dim_feedforward=None,
        encoder_conformer=True,
        encoder_depthwise_conv=True,
        encoder_conv_filter_size=1024,
        decoder_hidden=256,
        decoder_head=2,
        decoder_layers=4,
        decoder_dropout=0.1,
        decoder_kernel_sizes=[17, 21, 9, 13],
        decoder_dim_feedforward=None,
        decoder_conformer=True,
        decoder_depthwise_conv=True,
        decoder_conv_filter_size=1024,
        valid_nexamples=10,
        valid_example_directory=None,
        variance_early_stopping="none",
        variance_early_stopping_patience=4,
        variance_early_stopping_directory="variance_encoders",
        num_workers=num_cpus,
        cache_path=None,
        priors_gmm=False,
        priors_gmm_max_components=5,
        priors_gmm_min_samples_per_component=20,
        priors_gmm_reg_covar=1e-3,
        priors_gmm_logs=[0, 1, 2, 3],
        dvector_gmm=False,
        fastdiff_model=None,
        fastdiff_schedule=[0, 1],
        fastdiff_schedule_start=0,
        fastdiff_schedule_end=20,
        sort_data_by_length=False,
        fastdiff_variances=True,
        fastdiff_speakers=False,
        fastdiff_speakers_loss_weight=1,
        # additions for dataset-free construction (synthetic weights / benchmarks)
        stats=None,
        phone2id=None,
        fastdiff_head=False,
    ):
        super().__init__()
        self.lr = lr
        self.warmup_steps = warmup_steps
        self.num_workers = num_workers
        self.valid_nexamples = valid_nexamples
        self.valid_example_directory = valid_example_directory
        self.batch_size = batch_size
        self.save_hyperparameters(ignore=[
            "train_ds", "valid_ds", "train_ds_kwargs", "valid_ds_kwargs", "valid_nexamples",
            "valid_example_directory", "batch_size", "variance_early_stopping_directory", "num_workers",
            "fastdiff_model", "stats", "phone2id", "fastdiff_head"])
        hp = self.hparams

        # -- what the reference would reject or what is outside the hot path ------------------
        if fastdiff_model is not None or fastdiff_speakers:
            raise NotImplementedError("joint FastDiff vocoder / speaker generator")
        if "dvector" not in speaker_type:
            raise NotImplementedError("only d-vector speakers work at the reference HEAD (SURVEY 8, quirk 2)")
        if speaker_embedding_every_layer or prior_embedding_every_layer:
            raise NotImplementedError("*_embedding_every_layer (TypeError in the reference, quirk 4)")
        if not (encoder_conformer and decoder_conformer):
            raise NotImplementedError("plain TransformerEncoderLayer stacks")
        if encoder_hidden != decoder_hidden:
            raise NotImplementedError("encoder_hidden != decoder_hidden")
        self.fastdiff_model = None
        self.fastdiff_speaker_generator = None

        # -- datasets (reference :167-228: raw alignment datasets are wrapped in TTSDataset, behind the pickle cache;
        #    built datasets pass through) and the metadata derived from them (:236-245)
        train_ds, valid_ds = boundary.build_datasets(
            train_ds, valid_ds, train_ds_kwargs, valid_ds_kwargs, cache_path,
            dict(speaker_type=speaker_type, min_length=min_length, max_length=max_length,
                 augment_duration=augment_duration, variances=variances, variance_levels=variance_levels,
                 variance_transforms=variance_transforms, priors=priors, n_mels=n_mels, sampling_rate=sampling_rate,
                 n_fft=n_fft, win_length=win_length, hop_length=hop_length))
        if train_ds is not None:
            self.train_ds = train_ds
            self.stats = train_ds.stats
            self.phone2id = train_ds.phone2id
            if "dvector" in getattr(train_ds, "speaker_type", speaker_type):
                self.speaker2dvector = getattr(train_ds, "speaker2dvector", {})
            if getattr(train_ds, "speaker_type", speaker_type) == "id":
                self.speaker2id = getattr(train_ds, "speaker2id", {})
        if valid_ds is not None:
            self.valid_ds = valid_ds
        if stats is not None:
            self.stats = stats
        if phone2id is not None:
            self.phone2id = phone2id

        if hasattr(self, "phone2id"):
            self.phone_embedding = nn.Embedding(len(self.phone2id), hp.encoder_hidden, padding_idx=0)

        def stack(hidden, head, nlayers, ksizes, fsz, dropout, depthwise):
            return TransformerStack([
                ConformerEncoderLayer(hidden, head, conv_in=hidden, conv_filter_size=fsz,
                                      conv_kernel=(ksizes[i], 1), batch_first=True, dropout=dropout,
                                      conv_depthwise=depthwise)
                for i in range(nlayers)])

        self.encoder = stack(hp.encoder_hidden, hp.encoder_head, hp.encoder_layers, hp.encoder_kernel_sizes,
                             hp.encoder_conv_filter_size, hp.encoder_dropout, hp.encoder_depthwise_conv)
        self.positional_encoding = PositionalEncoding(hp.encoder_hidden, dropout=hp.encoder_dropout)
        if hasattr(self, "stats"):
            self.variance_adaptor = self._make_variance_adaptor()
        self.decoder = stack(hp.decoder_hidden, hp.decoder_head, hp.decoder_layers, hp.decoder_kernel_sizes,
                             hp.decoder_conv_filter_size, hp.decoder_dropout, hp.decoder_depthwise_conv)
        self.linear = nn.Linear(hp.decoder_hidden, hp.n_mels)
        self._mel_pack = _PackCache()
        if fastdiff_head:  # same shape as the reference's head (:393-402); feeds only result["fastdiff_var"]
            self.fastdiff_linear = nn.Sequential(nn.Linear(hp.decoder_hidden, hp.decoder_hidden),
                                                 nn.Linear(hp.decoder_hidden, hp.n_mels))
        if hasattr(self, "stats"):  # reference :416-424
            self.prior_embeddings = nn.ModuleDict({
                prior: PriorEmbedding(hp.encoder_hidden, hp.variance_nbins, self.stats[f"{prior}_prior"])
                for prior in hp.priors})
        self.speaker_embedding = SpeakerEmbedding(hp.encoder_hidden, hp.speaker_type)

        loss_weights = {"mel": hp.mel_loss_weight, "duration": hp.duration_loss_weight,
                        "speakers": hp.fastdiff_speakers_loss_weight}
        for i, var in enumerate(hp.variances):
            loss_weights[var] = hp.variance_loss_weights[i]
        self.loss = FastSpeech2Loss(hp.variances, hp.variance_levels, hp.variance_transforms, hp.variance_losses,
                                    hp.mel_loss, hp.duration_loss, hp.duration_stochastic, self._max_frames(),
                                    loss_weights)

    # ------------------------------------------------------------------------------------
    compute_mode = "fp32"

    def set_compute_mode(self, mode):
        """"fp32": tcgen05 GEMM/attention on bf16 hi/lo split operands, 3 passes (fp32 parity, mel within 1e-3);
        "bf16": single-pass bf16 operands, fp32 accumulate/residual/LayerNorm/softmax (mel within 1e-2);
        "simt": the exact-fp32 CUDA-core kernels (ground truth on device, any shape)."""
        if mode not in COMPUTE_MODES:
            raise ValueError(f"compute_mode must be one of {COMPUTE_MODES}")
        for m in self.modules():
            if hasattr(type(m), "compute_mode"):
                m.compute_mode = mode
        return self

    def _max_frames(self):
        hp = self.hparams
        return hp.max_length * hp.sampling_rate / hp.hop_length  # float, 2756.25 by default

    def _make_variance_adaptor(self):
        hp = self.hparams
        if hp.fastdiff_variances:  # reference :302-320
            from .fastdiff_variances import FastDiffVarianceAdaptor

            return FastDiffVarianceAdaptor(
                self.stats, hp.variances, hp.variance_nlayers, hp.variance_kernel_size, hp.variance_dropout,
                hp.variance_filter_size, hp.variance_nbins, hp.variance_depthwise_conv, hp.duration_nlayers,
                hp.duration_kernel_size, hp.duration_dropout, hp.duration_filter_size, hp.duration_depthwise_conv,
                hp.encoder_hidden, self._max_frames()).to(self.device)
        return VarianceAdaptor(
            self.stats, hp.variances, hp.variance_levels, hp.variance_transforms, hp.variance_nlayers,
            hp.variance_kernel_size, hp.variance_dropout, hp.variance_filter_size, hp.variance_nbins,
            hp.variance_depthwise_conv, hp.duration_nlayers, hp.duration_stochastic, hp.duration_kernel_size,
            hp.duration_dropout, hp.duration_filter_size, hp.duration_depthwise_conv, hp.encoder_hidden,
            self._max_frames()).to(self.device)

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .float(): parameters move, every captured graph and the cached parameter list is stale
        self.__dict__.pop("_graph_params", None)
        self.__dict__.pop("_graphs", None)
        return super()._apply(fn, *args, **kwargs)

    # -- checkpoint hooks (reference :530-634) -----------------------------------------------
    def on_load_checkpoint(self, checkpoint):
        self.__dict__.pop("_graph_params", None)
        self.stats = checkpoint["stats"]
        if not hasattr(self, "variance_adaptor"):
            self.variance_adaptor = self._make_variance_adaptor()
        self.phone2id = checkpoint["phone2id"]
        if not hasattr(self, "phone_embedding"):
            self.phone_embedding = nn.Embedding(len(self.phone2id), self.hparams.encoder_hidden, padding_idx=0)
        if not hasattr(self, "prior_embeddings"):  # model built without stats (reference :555-563)
            hp = self.hparams
            self.prior_embeddings = nn.ModuleDict({
                prior: PriorEmbedding(hp.encoder_hidden, hp.variance_nbins, self.stats[f"{prior}_prior"])
                for prior in hp.priors}).to(self.device)
        for key in ("speaker2dvector", "speaker2priors", "speaker_gmms", "dvector_gmms"):
            if key in checkpoint:
                setattr(self, key, checkpoint[key])
        state_dict = checkpoint["state_dict"]
        if any(k.startswith("fastdiff_linear.") for k in state_dict) and not hasattr(self, "fastdiff_linear"):
            d = self.hparams.decoder_hidden
            self.fastdiff_linear = nn.Sequential(nn.Linear(d, d), nn.Linear(d, self.hparams.n_mels))
        model_state = self.state_dict()
        changed = False
        for k in list(state_dict):
            if k in model_state:
                if state_dict[k].shape != model_state[k].shape:
                    print(f"Skip loading parameter: {k}, required shape: {model_state[k].shape}, "
                          f"loaded shape: {state_dict[k].shape}")
                    state_dict[k] = model_state[k]
                    changed = True
            else:
                print(f"Dropping parameter {k}")
                changed = True
        if changed:
            checkpoint.pop("optimizer_states", None)

    def on_save_checkpoint(self, checkpoint):
        checkpoint["stats"] = self.stats
        checkpoint["phone2id"] = self.phone2id
        for key in ("speaker2id", "speaker2dvector", "speaker2priors", "speaker_gmms", "dvector_gmms"):
            if hasattr(self, key):
                checkpoint[key] = getattr(self, key)

    # -- THE hot path (reference :636-784) ---------------------------------------------------
    def forward(self, targets, inference=False, force=None, control=None):
        """``force`` (optional, not in the reference): {"duration_rounded", "bucket_idx":{var:..},
        "want_idx": bool} teacher-forces / reports the two discrete decisions for parity runs."""
        dev = self.device
        if dev.type != "cuda":
            raise ops._lib.Lfs2Error("FastSpeech2.forward needs the model on a CUDA device: there is no CPU path")
        hp = self.hparams
        if hp.fastdiff_variances:
            return self._forward_fastdiff(targets, inference, force or {})
        if not inference and self.training and torch.is_grad_enabled() and not force and not control:
            return self._forward_train(targets)
        if inference and self.length_buckets > 1 and targets["phones"].shape[0] > 1:
            return self._forward_bucketed(targets, control, force)
        if inference and self.cuda_graphs and not force and not control and ops.PROFILE is None \
                and not self._skips_pad_rows(inference):
            return self._forward_graphed(targets)
        st = self._encode_stage(targets, inference, force, control)
        return self._decode_stage(st, targets, inference, force, control)

    def _encode_stage(self, targets, inference, force, control=None):
        """front end, encoder, duration predictor and the durations used (reference :639-686 + model.py:249-309)"""
        dev, hp = self.device, self.hparams
        phones = targets["phones"].to(dev, non_blocking=True).contiguous()
        speakers = targets["speaker"].to(dev, dtype=torch.float32, non_blocking=True).contiguous()
        spk = self.speaker_embedding.project(speakers)                       # (B, d)
        pe = self.positional_encoding.pe
        if self.training and hp.encoder_dropout > 0:
            raise NotImplementedError("dropout in training mode outside the train step (use model.train() with grad "
                                      "enabled for the train path, or .eval())")
        output, src_mask = ops.embed_pe_spk(phones, self.phone_embedding.weight, pe, spk)
        limit = None
        if self._skips_pad_rows(inference):
            limit = (ops.mask_lengths(src_mask), self._halos()[0])
        output = self.encoder(output, src_key_padding_mask=src_mask, row_limit=limit)
        if len(hp.priors):  # per-utterance prior embeddings, broadcast over the phones (reference :687-692)
            zero_pe = torch.zeros(1, max(output.shape[1], 1), output.shape[2], device=dev)
            for prior in hp.priors:
                term, _ = self.prior_embeddings[prior].term(targets[f"priors_{prior}"])
                ops.add_pe_spk_(output, zero_pe, term)
        st = self.variance_adaptor.durations(output, src_mask, targets, inference=inference, force=force,
                                             control=control)
        st.update(enc=st["x_phone"], src_mask=src_mask, spk=spk)
        return st

    def _decode_stage(self, st, targets, inference, force, control, scan=None, frames=None):
        """LengthRegulator, variance encoders, decoder, mel Linear, result dict (reference model.py:311-341, :703-784)"""
        dev, hp = self.device, self.hparams
        pe, spk, src_mask = self.positional_encoding.pe, st["spk"], st["src_mask"]
        self.variance_adaptor.need_out_val = hasattr(self, "fastdiff_linear")
        tc_decoder = (self.compute_mode != "simt" and hp.decoder_hidden % 32 == 0 and hp.n_mels % 16 == 0
                      and self.decoder.tc_capable(hp.decoder_hidden) and not (self.training and hp.decoder_dropout > 0))
        # tensor-core decoder: its input is only ever read as Planes -- the last variance embedding add, the positional /
        # speaker add and the plane split are one kernel (ops.decoder_input_planes), no fp32 tensor in between
        tail = (pe, spk, self.decoder.wants_f16_plane(hp.decoder_hidden)) if tc_decoder else None
        variance_output = self.variance_adaptor.expand(st["enc"], st, targets, inference=inference, force=force,
                                                       control=control, scan=scan, frames=frames, tail=tail)
        output = variance_output["x_planes"] if tail is not None else ops.add_pe_spk_(variance_output["x"], pe, spk)
        tgt_mask = variance_output["tgt_mask"]
        if self._skips_pad_rows(inference):
            # PAD rows farther than the decoder's conv halo past an utterance's end: never computed, never written
            lens = ops.mask_lengths(tgt_mask)
            output = self.decoder(output, src_key_padding_mask=tgt_mask, return_planes=True,
                                  row_limit=(lens, sum(layer.halo() for layer in self.decoder.layers)))
            wmel = self._mel_pack.get([self.linear.weight], lambda: ops.split_bf16(self.linear.weight.detach().contiguous()))
            mel = ops.gemm_tc(output, wmel, self.linear.bias, npass=3 if self.compute_mode == "fp32" else 1,
                              tag="mel_linear", row_limit=(lens, 0))
            ops.zero_masked_rows_(mel, tgt_mask)
        elif self.compute_mode != "simt" and hp.decoder_hidden % 32 == 0 and hp.n_mels % 16 == 0:
            output = self.decoder(output, src_key_padding_mask=tgt_mask, return_planes=True)
            wmel = self._mel_pack.get([self.linear.weight], lambda: ops.split_bf16(self.linear.weight.detach().contiguous()))
            mel = ops.gemm_tc(output, wmel, self.linear.bias, npass=3 if self.compute_mode == "fp32" else 1,
                              tag="mel_linear")
        else:
            output = self.decoder(output, src_key_padding_mask=tgt_mask)
            mel = ops.linear(output, self.linear.weight, self.linear.bias, tag="mel_linear")

        result = {
            "mel": mel,
            "duration_prediction": variance_output["duration_prediction"],
            "duration_rounded": variance_output["duration_rounded"],
            "src_mask": src_mask,
            "tgt_mask": tgt_mask,
            "frame_lengths": variance_output.get("frame_lengths"),
            "frame_lengths_host": variance_output.get("frame_lengths_host"),
        }
        if hasattr(self, "fastdiff_linear") and variance_output["out"] is not None:
            zero_pe = torch.zeros(1, mel.shape[1], hp.decoder_hidden, device=dev)
            h = ops.add_pe_spk_(variance_output["out"], zero_pe, spk)
            h = ops.linear(h, self.fastdiff_linear[0].weight, self.fastdiff_linear[0].bias)
            h = ops.linear(h, self.fastdiff_linear[1].weight, self.fastdiff_linear[1].bias)
            result["fastdiff_var"] = h * 0.1
        for var in hp.variances:
            result[f"variances_{var}"] = variance_output[f"variances_{var}"]
            if f"_bucket_{var}" in variance_output:
                result[f"_bucket_{var}"] = variance_output[f"_bucket_{var}"]
        return result

    # -- FastDiff variance adaptor (reference fastdiff_variances.py, fastspeech2.py:302-320, 769-776) ----
    def _forward_fastdiff(self, targets, inference, force):
        """Same path with the diffusion adaptor between encoder and decoder.  ``force`` may carry the random draws of a
        parity run: "noise" (list of tensors in the reference's draw order), "steps" ({name: (B) int64}), "jitter"
        ((B, Tp)), "N" (reverse steps, default 4) and "duration_rounded"."""
        dev, hp = self.device, self.hparams
        if self.training and torch.is_grad_enabled():
            raise NotImplementedError("training the FastDiff variance adaptor needs backward kernels that are not written; "
                                      "inference and the teacher-forced forward (eval / no_grad) run on the CUDA path")
        if self.skip_pad_rows or self.length_buckets > 1:
            raise NotImplementedError("skip_pad_rows / length_buckets with the FastDiff variance adaptor")
        phones = targets["phones"].to(dev, non_blocking=True).contiguous()
        speakers = targets["speaker"].to(dev, dtype=torch.float32, non_blocking=True).contiguous()
        spk = self.speaker_embedding.project(speakers)
        pe = self.positional_encoding.pe
        output, src_mask = ops.embed_pe_spk(phones, self.phone_embedding.weight, pe, spk)
        output = self.encoder(output, src_key_padding_mask=src_mask)
        if len(hp.priors):
            zero_pe = torch.zeros(1, max(output.shape[1], 1), output.shape[2], device=dev)
            for prior in hp.priors:
                term, _ = self.prior_embeddings[prior].term(targets[f"priors_{prior}"])
                ops.add_pe_spk_(output, zero_pe, term)
        noise = list(force["noise"]) if force.get("noise") is not None else None
        vo = self.variance_adaptor(output, src_mask, targets, inference=inference, N=force.get("N", 4), noise=noise,
                                   steps=force.get("steps"), jitter=force.get("jitter"),
                                   force={k: v for k, v in force.items() if k == "duration_rounded"})
        tgt_mask = vo["tgt_mask"]
        output = ops.add_pe_spk_(vo["x"], pe, spk)
        if self.compute_mode != "simt" and hp.decoder_hidden % 32 == 0 and hp.n_mels % 16 == 0:
            outp = self.decoder(output, src_key_padding_mask=tgt_mask, return_planes=True)
            wmel = self._mel_pack.get([self.linear.weight], lambda: ops.split_bf16(self.linear.weight.detach().contiguous()))
            mel = ops.gemm_tc(outp, wmel, self.linear.bias, npass=3 if self.compute_mode == "fp32" else 1, tag="mel_linear")
        else:
            mel = ops.linear(self.decoder(output, src_key_padding_mask=tgt_mask), self.linear.weight, self.linear.bias,
                             tag="mel_linear")
        result = {"mel": mel, "duration_prediction": vo["duration_prediction"], "duration_rounded": vo["duration_rounded"],
                  "src_mask": src_mask, "tgt_mask": tgt_mask}
        if hasattr(self, "fastdiff_linear") and vo["out"] is not None:
            zero_pe = torch.zeros(1, mel.shape[1], hp.decoder_hidden, device=dev)
            h = ops.add_pe_spk_(vo["out"], zero_pe, spk)
            h = ops.linear(h, self.fastdiff_linear[0].weight, self.fastdiff_linear[0].bias)
            h = ops.linear(h, self.fastdiff_linear[1].weight, self.fastdiff_linear[1].bias)
            result["fastdiff_var"] = h * 0.1
        for var in hp.variances:
            result[f"variances_{var}"] = vo[f"variances_{var}"]
            result[f"variances_{var}_z"] = vo[f"variances_{var}_z"]
        result["duration_z"] = vo["duration_z"]
        return result

    # -- PAD-row skipping synthesis ("drop PAD-row compute where provably unobservable") ----------
    # Off by default: the default path computes every PAD row like the reference does.  When on, inference runs
    # every encoder / decoder kernel only over the 128-row tiles that start before an utterance's end + the summed
    # conv half-widths downstream (`_halos`): PAD rows are never attention keys, so rows farther out cannot reach a
    # valid row, and the depthwise convs read the rows past the last kept tile as zeros.  Valid mel frames (and all
    # predictions) are bit-identical to the default path; mel frames masked by tgt_mask come back as zeros.
    skip_pad_rows = False

    def _skips_pad_rows(self, inference):
        if not (self.skip_pad_rows and inference) or self.length_buckets > 1:
            return False
        hp = self.hparams
        if self.compute_mode == "simt" or hp.n_mels % 16 != 0 or any(lv != "frame" for lv in hp.variance_levels) \
                or not self.encoder.supports_row_limit(hp.encoder_hidden) \
                or not self.decoder.supports_row_limit(hp.decoder_hidden):
            raise NotImplementedError("skip_pad_rows needs tensor-core FFTBlocks with a row-limited path (d = 256 with "
                                      "head_dim 128 and the fused depthwise FFN, or head_dim 256 / 384 with the depthwise "
                                      "FFN), frame-level variances, compute mode fp32 or bf16")
        return True

    # -- length-bucketed synthesis ---------------------------------------------------
    length_buckets = 1
    bucket_graphs = True   # replay each bucket's kernel sequence as a CUDA graph once its shape has been seen twice
    max_graphs = 32

    def _graphs_validate(self):
        """graphs bake in the addresses of the packed weights: drop them when any parameter changed"""
        plist = self.__dict__.get("_graph_params")
        if plist is None:  # walking the module tree costs ~1 ms; the Parameter objects only change with the module set
            plist = self.__dict__["_graph_params"] = list(self.parameters())
        sig = (ops.WEIGHTS_EPOCH, self.compute_mode, sum(p._version for p in plist),
               sum(p.data_ptr() for p in plist) & 0xFFFFFFFFFFFF)
        if getattr(self, "_graph_sig", None) != sig:
            self._graphs = {}
            self._graph_sig = sig

    def _graphed(self, key, fn, inputs):
        """Run fn(*inputs) eagerly the first time `key` is seen (this also builds the weight packs), capture it
        into a CUDA graph the second time, replay afterwards: a bucket is ~90 short launches and the host
        (Python + ctypes, ~30 us per launch) is otherwise the bottleneck.  Returns (outputs, replayed);
        replayed outputs live in the graph's private pool and are overwritten by the next replay."""
        cache = self.__dict__.setdefault("_graphs", {})
        entry = cache.get(key)
        if entry is None:
            if len(cache) >= self.max_graphs:
                cache.pop(next(iter(cache)))
            # `gen` distinguishes this entry from an evicted one with the same key: graphs captured downstream of it
            # (the decoder stage reads the encoder graph's static output buffers) carry it in THEIR key
            self.__dict__["_graph_gen"] = self.__dict__.get("_graph_gen", 0) + 1
            cache[key] = {"seen": 1, "gen": self.__dict__["_graph_gen"]}
            return fn(*inputs), False
        if "graph" not in entry:
            static_in = [torch.empty_like(x, device=self.device) for x in inputs]
            for sbuf, x in zip(static_in, inputs):
                sbuf.copy_(x, non_blocking=True)
            calls0 = ops._lib.CALLS
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = fn(*static_in)
            entry.update(graph=graph, static_in=static_in, out=out, launches=ops._lib.CALLS - calls0)
            ops._lib.CALLS = calls0
        for sbuf, x in zip(entry["static_in"], inputs):
            sbuf.copy_(x, non_blocking=True)
        entry["graph"].replay()
        ops._lib.CALLS += entry["launches"]
        return entry["out"], True

    # -- whole-call CUDA graphs (small batches are launch-bound: ~100 short kernels, ~15 us of Python + ctypes each) -----
    # `model.cuda_graphs = True`: an inference call is split at its one host read-back (the LengthRegulator's frame
    # count) into an encoder-side and a decoder-side kernel sequence; each is captured the second time its shape is seen
    # and replayed afterwards (same machinery and invalidation rules as the bucketed path).  Results are copied out of
    # the graphs' static buffers, so they stay valid across calls.  Bit-identical to the eager path.
    cuda_graphs = False

    def _forward_graphed(self, targets):
        dev = self.device
        self._graphs_validate()
        phones = targets["phones"].to(dev, non_blocking=True).contiguous()
        speaker = targets["speaker"].to(dev, dtype=torch.float32, non_blocking=True).contiguous()

        def enc_fn(ph, sp):
            st = self._encode_stage({"phones": ph, "speaker": sp}, True, None, None)
            st["scan"] = ops.length_regulate_scan(st["duration_rounded"], st["enc"].shape[:2])
            return st

        key = ("EW", tuple(phones.shape), self.compute_mode)
        st, replayed = self._graphed(key, enc_fn, [phones, speaker])
        longest = int(st["scan"][2].item())  # the single device->host sync of the path
        l = min(longest, int(self.variance_adaptor.max_length))

        def dec_fn(st=st, l=l):
            return self._decode_stage(st, None, True, None, None, scan=st["scan"], frames=(l, l))

        if replayed:  # the encoder stage's outputs live at fixed addresses: the decoder graph can bake them in
            r, dec_replayed = self._graphed(("DW", key, self._graphs[key]["gen"], l), dec_fn, [])
        else:
            r, dec_replayed = dec_fn(), False
        if replayed or dec_replayed:  # static buffers are overwritten by the next replay
            r = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in r.items()}
        return r

    def _halos(self):
        """(encoder side, decoder side): rows beyond an utterance's end that can still influence its valid
        rows -- the summed conv half-widths of everything downstream (attention never reads PAD keys)."""
        va = self.variance_adaptor
        h_enc = sum(layer.halo() for layer in self.encoder.layers) + va.duration_predictor.halo()
        h_dec = max([sum(layer.halo() for layer in self.decoder.layers)] +
                    [va.encoders[v].predictor.halo() for v in va.variances])
        return h_enc, h_dec

    def _forward_bucketed(self, targets, control, force=None):
        """Synthesis of a ragged batch as `length_buckets` length-sorted sub-batches, each padded only to ITS
        longest utterance plus the conv halo.  Exactness: PAD rows are never attention keys, so they reach
        valid rows only through the FFN / predictor convolutions; every row within the summed half-widths
        (`_halos`) of an utterance's end is kept (and holds exactly the value the reference computes there:
        PE[t] + speaker term, ...), rows farther out are provably unobservable on valid rows.  Valid
        positions therefore equal the un-bucketed result; positions the reference's consumers mask with
        tgt_mask (generator.py:164) are returned as zeros instead of the reference's PAD-row values."""
        dev, hp = self.device, self.hparams
        phones_all = targets["phones"]
        speaker_all = targets["speaker"]
        bsz, tp = phones_all.shape
        nz = (phones_all != 0)
        lengths = (nz * torch.arange(1, tp + 1, device=phones_all.device)).amax(1).tolist()  # last valid phone + 1
        order = sorted(range(bsz), key=lambda i: -lengths[i])
        ngroups = min(self.length_buckets, bsz)
        per = (bsz + ngroups - 1) // ngroups
        h_enc, h_dec = self._halos()
        cap = int(self.variance_adaptor.max_length)
        # phase 1, per bucket: encoder + durations (the encoder side may be cut at the bucket's longest utterance
        # + halo because the full tensor's end, where the reference zero-pads, is known: tp)
        use_graphs = self.bucket_graphs and not force and not control and ops.PROFILE is None
        if use_graphs:
            self._graphs_validate()
        stages = []
        for g0 in range(0, bsz, per):
            idx = order[g0:g0 + per]
            tp_g = min(tp, max(lengths[i] for i in idx) + h_enc)
            it = torch.tensor(idx, device=phones_all.device)
            sub = {"phones": phones_all[it][:, :tp_g].contiguous(), "speaker": speaker_all[it].contiguous()}
            f = None
            if force:  # parity runs: the reference's discrete decisions, sliced to this bucket
                f = {"want_idx": force.get("want_idx", False)}
                if "duration_rounded" in force:
                    f["duration_rounded"] = force["duration_rounded"][it.to(force["duration_rounded"].device)][:, :tp_g]
                if "bucket_idx" in force:
                    f["bucket_idx"] = {v: t[it.to(t.device)] for v, t in force["bucket_idx"].items()}

            def enc_fn(phones, speaker, f=f):
                st = self._encode_stage({"phones": phones, "speaker": speaker}, True, f, control)
                st["scan"] = ops.length_regulate_scan(st["duration_rounded"], st["enc"].shape[:2])
                return st

            # one graph per bucket ORDINAL: two buckets of equal shape must not share static output buffers, since
            # every encoder stage of the batch runs before the first decoder stage
            key = ("E", g0, len(idx), tp_g, self.compute_mode)
            if use_graphs:
                st, replayed = self._graphed(key, enc_fn, [sub["phones"], sub["speaker"]])
                ekey = (key, self._graphs[key]["gen"]) if replayed else None
            else:
                st, ekey = enc_fn(sub["phones"], sub["speaker"]), None
            stages.append((idx, tp_g, f, st, ekey))
        # the decoder side needs the global frame count first (where the reference's tensor ends): ONE read-back
        longest = torch.cat([st["scan"][2] for _, _, _, st, _ in stages]).tolist()
        l_glob = min(max(longest), cap)
        parts = []
        for (idx, tp_g, f, st, ekey), lg in zip(stages, longest):
            cap_g = min(lg, cap)
            frames = (min(cap_g + h_dec, l_glob), cap_g)

            def dec_fn(st=st, f=f, frames=frames):
                return self._decode_stage(st, None, True, f, control, scan=st["scan"], frames=frames)

            if ekey is not None:  # the encoder stage was a graph replay: its outputs live at fixed addresses
                r, _ = self._graphed(("D", ekey, frames), dec_fn, [])
            else:
                r = dec_fn()
            parts.append((idx, tp_g, r))
        n_mels = hp.n_mels
        out = {
            "mel": torch.zeros(bsz, l_glob, n_mels, device=dev),
            "duration_prediction": torch.zeros(bsz, tp, device=dev),
            "duration_rounded": torch.zeros(bsz, tp, device=dev, dtype=torch.int32),
            "src_mask": phones_all.to(dev) == 0,
            "tgt_mask": torch.ones(bsz, l_glob, device=dev, dtype=torch.bool),
        }
        for v in hp.variances:
            out[f"variances_{v}"] = torch.zeros(bsz, l_glob, device=dev)
            if any(f"_bucket_{v}" in r for _, _, r in parts):
                out[f"_bucket_{v}"] = torch.zeros(bsz, l_glob, device=dev, dtype=torch.int64)
        if any("fastdiff_var" in r for _, _, r in parts):
            out["fastdiff_var"] = torch.zeros(bsz, l_glob, n_mels, device=dev)
        for idx, tp_g, r in parts:
            it = torch.tensor(idx, device=dev)
            w = min(r["mel"].shape[1], l_glob)
            for key in ("mel", "tgt_mask", "fastdiff_var", *[f"variances_{v}" for v in hp.variances],
                        *[f"_bucket_{v}" for v in hp.variances]):
                if key in r and key in out:
                    out[key][it, :w] = r[key][:, :w]
            out["duration_prediction"][it, :tp_g] = r["duration_prediction"]
            out["duration_rounded"][it, :tp_g] = r["duration_rounded"].to(torch.int32)
        return out

    # -- train step: forward with saved activations + hand-written backward (training.py) -----
    # train_length_buckets = n > 1: the train step of a ragged batch runs as n length-sorted sub-batches, each padded to
    # its own longest utterance + the conv halo (training.forward_train_bucketed); losses and gradients equal the
    # un-bucketed step up to fp32 summation order; result positions past a bucket's own tensor (masked) are zeros.  Off by default.
    train_length_buckets = 1

    def _forward_train(self, targets):
        """Teacher-forced forward (reference :636-784 with inference=False) whose outputs are connected
        to autograd through ONE Function; its backward runs the liblfs2.so gradient kernels and
        accumulates into p.grad (see training.py)."""
        if not hasattr(self, "_grad_anchor") or self._grad_anchor.device != self.device:
            self._grad_anchor = torch.zeros((), device=self.device, requires_grad=True)
        keys = ["mel", "duration_prediction"] + [f"variances_{v}" for v in self.hparams.variances]
        outs = training.ForwardTrainFn.apply(self._grad_anchor, self, targets, keys)
        result = dict(self._last_train_result)
        self._last_train_result = None
        result.update(zip(keys, outs))
        return result

    def flatten_parameters(self):
        """Re-home every trainable parameter (and its gradient) as a view of ONE flat fp32 buffer
        (each tensor 128-byte aligned): the fused AdamW kernel and the single NCCL all-reduce of the
        gradient path work on these buffers.  Call after .to(device); idempotent."""
        params = [p for p in self.parameters() if p.requires_grad]
        if getattr(self, "_flat_ids", None) == [id(p) for p in params] and self._flat_param.device == self.device:
            return self._flat_param, self._flat_grad
        offs, n = [], 0
        for p in params:
            offs.append(n)
            n += (p.numel() + 31) // 32 * 32
        dev = self.device
        flat_p = torch.zeros(n, device=dev, dtype=torch.float32)
        flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        with torch.no_grad():
            for p, o in zip(params, offs):
                v = flat_p[o:o + p.numel()].view(p.shape)
                v.copy_(p.data)
                p.data = v
                p.grad = flat_g[o:o + p.numel()].view(p.shape)
        self._flat_param, self._flat_grad = flat_p, flat_g
        self._flat_ids = [id(p) for p in params]
        self.__dict__["_flat_layout"] = (params, offs)
        ops.WEIGHTS_EPOCH += 1
        return flat_p, flat_g

    def rehome_gradients(self):
        """Make every p.grad a view of the flat gradient buffer again.  ``zero_grad(set_to_none=True)`` (torch's and
        Lightning's default) drops the views; a gradient tensor that is not a view would never reach the fused AdamW
        or the all-reduce.  A stray gradient that already holds values is added into its slot first."""
        layout = self.__dict__.get("_flat_layout")
        if layout is None:
            return
        flat_g = self._flat_grad
        base = flat_g.data_ptr()
        for p, o in zip(*layout):
            g = p.grad
            if g is not None and g.data_ptr() == base + 4 * o:
                continue
            view = flat_g[o:o + p.numel()].view(p.shape)
            if g is not None:
                with torch.no_grad():
                    view.add_(g.to(view.dtype))
            p.grad = view

    def zero_grad(self, set_to_none=True):
        """With flat buffers the gradients are zeroed in place and stay views (set_to_none would orphan them)."""
        if self.__dict__.get("_flat_layout") is None:
            return super().zero_grad(set_to_none=set_to_none)
        self._flat_grad.zero_()
        self.rehome_gradients()

    def allreduce_gradients(self, group=None):
        """Data-parallel gradient exchange: ONE NCCL all-reduce (sum) over the flat gradient buffer
        (SURVEY 8e); the 1/world_size average is folded into FusedAdamW.  Returns the world size."""
        from ..sharding import allreduce_sum_

        _, flat_g = self.flatten_parameters()
        return allreduce_sum_(flat_g, group)

    # Under pl.Trainer: parameters never enter torch's autograd graph (the backward is hand-written and writes p.grad
    # itself), so torch's DistributedDataParallel wrapper -- Lightning's strategy="ddp" -- has no AccumulateGrad hooks to
    # ride on and MUST NOT wrap this module.  Data-parallel training = one process per GPU with a strategy that leaves
    # the module unwrapped, and this hook (Lightning calls it right after loss.backward()) doing the one all-reduce.
    allreduce_in_hook = False

    def on_after_backward(self):
        if self.allreduce_in_hook:
            world = self.allreduce_gradients()
            if getattr(self, "optimizer", None) is not None and hasattr(self.optimizer, "grad_scale"):
                self.optimizer.grad_scale = 1.0 / world

    # -- train / validation steps (reference :786-807) ---------------------------------------
    def training_step(self, batch, batch_idx, optimizer_idx=0):
        result = self(batch, optimizer_idx)
        losses = self.loss(result, batch)
        if getattr(self, "log_losses", True):  # one D2H copy for all values (the reference does one .item() each)
            vals = self.loss.last_buffer.tolist()
            names = list(self.hparams.variances) + ["mel", "duration", "total"]
            self.log_dict({f"train/{k}_loss": v for k, v in zip(names, vals)}, batch_size=self.batch_size,
                          sync_dist=True)
        return losses["total"]

    def validation_step(self, batch, batch_idx):
        result = self(batch)
        losses = self.loss(result, batch)
        self.log_dict({f"eval/{k}_loss": v.item() for k, v in losses.items()}, batch_size=self.batch_size,
                      sync_dist=True)
        return self(batch, inference=True)

    def configure_optimizers(self):
        """AdamW(lr, betas (0.9, 0.98), eps 1e-8, wd 0.01) + NoamLR stepped every batch (reference
        :1166-1182).  On CUDA the optimizer is the fused flat-buffer AdamW kernel (same update rule,
        same param_groups / scheduler interface); on CPU (construction-time checks) torch's own."""
        if self.device.type == "cuda":
            self.optimizer = FusedAdamW(self, lr=self.hparams.lr, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.01)
        else:
            self.optimizer = torch.optim.AdamW(self.parameters(), lr=self.hparams.lr, betas=[0.9, 0.98], eps=1e-8,
                                               weight_decay=0.01)
        self.scheduler = NoamLR(self.optimizer, self.hparams.warmup_steps)
        return [self.optimizer], [{"scheduler": self.scheduler, "interval": "step"}]

    # -- argparse surface and dataloaders (reference :1184-1323; litfass/train.py:73-74, Trainer.fit) ----------
    @staticmethod
    def add_model_specific_args(parent_parser):
        return boundary.add_model_specific_args(parent_parser)

    @staticmethod
    def add_dataset_specific_args(parent_parser):
        return boundary.add_dataset_specific_args(parent_parser)

    def train_dataloader(self):
        if self.hparams.sort_data_by_length:
            self.train_ds.sort_by_duration()
        return boundary.dataloader(self.train_ds, self.batch_size, self.num_workers)

    def val_dataloader(self):
        return boundary.dataloader(self.valid_ds, self.batch_size, self.num_workers)


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction, same operation order) as
    ONE kernel over the model's flat parameter / gradient / moment buffers (28 bytes per parameter).
    ``grad_scale`` (1/world_size after the sum all-reduce) and optional global-norm clipping
    (Trainer(gradient_clip_val=...) in the reference's scripts/train.sh) are folded into the same pass,
    which also zeroes the gradient buffer for the next step.  The learning rate is read from
    ``param_groups[0]["lr"]``, so NoamLR drives it exactly like torch's optimizer."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.01, max_grad_norm=0.0):
        self.model = model
        self.flat_p, self.flat_g = model.flatten_parameters()
        params = [p for p in model.parameters() if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.step_count = 0
        self.grad_scale = 1.0
        self.max_grad_norm = max_grad_norm
        self._gnorm = torch.zeros(1, device=self.flat_p.device, dtype=torch.float32)

    def zero_grad(self, set_to_none=False):
        self.flat_g.zero_()
        self.model.rehome_gradients()

    def _slots(self):
        params, offs = self.model.__dict__["_flat_layout"]
        return [(o, p.numel(), p.shape) for p, o in zip(params, offs)]

    def state_dict(self):
        """torch.optim.AdamW's layout ({"state": {i: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups"}), so a
        checkpoint written here resumes under torch's AdamW (the reference's optimizer) and vice versa."""
        state = {}
        if self.step_count > 0:
            for i, (o, n, shape) in enumerate(self._slots()):
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": self.exp_avg[o:o + n].view(shape).clone(),
                            "exp_avg_sq": self.exp_avg_sq[o:o + n].view(shape).clone()}
        groups = [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]
        groups[0]["params"] = list(range(len(self._slots())))
        return {"state": state, "param_groups": groups, "lfs2_fused": {"step_count": self.step_count,
                                                                        "max_grad_norm": self.max_grad_norm}}

    def load_state_dict(self, sd):
        slots = self._slots()
        state = sd.get("state", {})
        if state and len(state) != len(slots):
            raise ValueError(f"optimizer state holds {len(state)} parameters, the model has {len(slots)} trainable ones")
        steps = set()
        with torch.no_grad():
            self.exp_avg.zero_()
            self.exp_avg_sq.zero_()
            for i, (o, n, shape) in enumerate(slots):
                st = state.get(i, state.get(str(i)))
                if st is None:
                    continue
                if tuple(st["exp_avg"].shape) != tuple(shape):
                    raise ValueError(f"optimizer state {i}: shape {tuple(st['exp_avg'].shape)} vs parameter {tuple(shape)}")
                self.exp_avg[o:o + n].view(shape).copy_(st["exp_avg"])
                self.exp_avg_sq[o:o + n].view(shape).copy_(st["exp_avg_sq"])
                steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): the fused kernel keeps one")
        self.step_count = steps.pop() if steps else int(sd.get("lfs2_fused", {}).get("step_count", 0))
        for g, saved in zip(self.param_groups, sd.get("param_groups", [])):
            g.update({k: v for k, v in saved.items() if k != "params"})
            g["betas"] = tuple(g["betas"])

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        self.model.rehome_gradients()
        self.step_count += 1
        gn = None
        if self.max_grad_norm > 0:
            self._gnorm.zero_()
            ops.sumsq_(self._gnorm, self.flat_g)
            gn = self._gnorm
        ops.adamw_step_(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, g["lr"], g["betas"][0], g["betas"][1],
                        g["eps"], g["weight_decay"], self.step_count, grad_scale=self.grad_scale,
                        max_norm=self.max_grad_norm, gnorm_sq=gn, zero_grad=True)
        ops.WEIGHTS_EPOCH += 1  # the kernel wrote through raw pointers: invalidate packed-weight caches
        return loss
