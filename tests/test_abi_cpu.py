"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol that
include/lfs2.h declares, and the host mirror keeps the reference's state_dict layout.
No compute calls are made (no GPU here)."""
import os
import re

import pytest
import torch

from lightningfastspeech2_b200 import _lib, configs
from lightningfastspeech2_b200.fastspeech2.fastspeech2 import FastSpeech2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lfs2.h")).read()
    return sorted(set(re.findall(r"LFS2_API\s+[\w\s\*]+?\b(lfs2_\w+)\s*\(", src)))


def test_header_and_binding_agree():
    declared = _declared()
    assert len(declared) >= 15
    bound = set(_lib.SIGNATURES) | set(_lib.RESTYPES) | {"lfs2_last_error"}
    assert set(declared) == bound, (set(declared) ^ bound)


def test_library_exports_every_declared_symbol():
    from lightningfastspeech2_b200 import build

    build.build()
    h = _lib.lib()
    for name in _declared():
        assert hasattr(h, name), name
    assert h.lfs2_version() >= 100
    assert isinstance(h.lfs2_last_error(), bytes)


def test_invalid_args_are_reported_not_thrown():
    h = _lib.lib()
    rc = h.lfs2_linear(None, None, None, None, 4, 4, 16, 0, None)
    assert rc == -1
    assert b"linear" in h.lfs2_last_error()
    with pytest.raises(_lib.Lfs2Error):
        _lib.check(rc, "lfs2_linear")


def test_no_cpu_fallback():
    from lightningfastspeech2_b200 import ops

    with pytest.raises(_lib.Lfs2Error):
        ops.linear(torch.zeros(4, 16), torch.zeros(8, 16), torch.zeros(8))


def _mk(preset):
    kw = configs.PRESETS[preset]
    stats = {v: {"min": -3.0, "max": 3.0, "mean": 0.0, "std": 1.0} for v in kw["variances"]}
    return FastSpeech2(stats=stats, phone2id={f"p{i}": i for i in range(80)}, fastdiff_head=True, num_workers=0, **kw)


@pytest.mark.parametrize("preset,golden", [("C1", "c1_infer"), ("C2", "c2_small_infer"),
                                           ("TINY_DW", "tiny_dw_infer"), ("TINY_DENSE", "tiny_dense_infer")])
def test_state_dict_layout_matches_reference(golden_dir, preset, golden):
    g = torch.load(os.path.join(golden_dir, golden + ".pt"), weights_only=False)
    sd = _mk(preset).state_dict()
    assert set(sd) == set(g["shapes"])
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(g["shapes"][k]), k


def test_param_counts_match_survey():
    for preset, count in (("C1", 24518737), ("C2", 7436881), ("C3", 75061329)):
        m = _mk(preset)
        n = sum(p.numel() for p in m.parameters()) - sum(p.numel() for p in m.fastdiff_linear.parameters())
        assert n == count, preset


def test_reference_default_kwargs_build_the_fastdiff_adaptor():
    """the reference's constructor defaults (fastdiff_variances=True, 3 frame-level variances) construct, with the
    reference's module tree and key names (fastspeech2.py:302-320)"""
    from lightningfastspeech2_b200.fastspeech2.fastdiff_variances import FastDiffVarianceAdaptor

    stats = {v: {"min": -3.0, "max": 3.0, "mean": 0.0, "std": 1.0} for v in ("pitch", "energy", "snr")}
    m = FastSpeech2(stats=stats, phone2id={"a": 0}, num_workers=0)
    assert isinstance(m.variance_adaptor, FastDiffVarianceAdaptor) and m.variance_adaptor.length_regulator.pad_to_multiple_of == 64
    keys = set(m.state_dict())
    assert {"variance_adaptor.duration_predictor.linear_in.weight", "variance_adaptor.duration_predictor.fc_t1.weight",
            "variance_adaptor.encoders.snr.predictor.linear_noise.bias", "variance_adaptor.encoders.pitch.bins",
            "variance_adaptor.encoders.energy.embedding.weight"} <= keys


def test_unsupported_configs_raise():
    with pytest.raises(NotImplementedError):
        FastSpeech2(stats={}, phone2id={"a": 0}, num_workers=0, **dict(configs.C2, speaker_type="id"))


@pytest.mark.gpu
def test_conv2_fold_is_exact_linear_map():
    """W_eff/b_eff (grouped 1x1 conv folded into the following pointwise conv by lfs2_fold_pw_fwd) reproduce
    conv2 of the reference layer, and two folds of the same weights give the same bits."""
    from lightningfastspeech2_b200.fastspeech2.model import ConformerEncoderLayer

    torch.manual_seed(0)
    layer = ConformerEncoderLayer(32, 2, conv_in=32, conv_filter_size=128, conv_kernel=(5, 1), batch_first=True,
                                  dropout=0.0, conv_depthwise=True)
    v = torch.randn(3, 128, 11)
    ref = layer.conv2(v)
    layer = layer.cuda()
    p = layer._build_pack()
    got = torch.einsum("nf,bft->bnt", p["w_eff"].cpu(), v) + p["b_eff"].cpu()[None, :, None]
    assert (ref - got).abs().max() < 1e-5
    q = layer._build_pack()
    assert torch.equal(p["w_eff"], q["w_eff"]) and torch.equal(p["b_eff"], q["b_eff"])


def test_noam_and_optimizer_recipe():
    m = _mk("TINY_DW")
    (opt,), (sched,) = m.configure_optimizers()
    g = opt.param_groups[0]
    assert g["betas"] == [0.9, 0.98] or tuple(g["betas"]) == (0.9, 0.98)
    assert g["eps"] == 1e-8 and g["weight_decay"] == 0.01
    assert sched["interval"] == "step"
    from oracle import fs2_oracle as O

    for step in (1, 10, 4000, 10000):
        from lightningfastspeech2_b200.fastspeech2.noam import noam_scale

        assert noam_scale(step, 4000) == pytest.approx(O.noam_scale(step, 4000))


def test_pad_row_skipping_host_logic():
    """model.skip_pad_rows: the halo the row limits are built from, and which configurations may use them"""
    m = _mk("C2").eval()
    # encoder FFN kernels 5,25,13,9 (+ duration predictor: 2 layers of k=3); decoder 17,21,9,13; predictors 5 x k=3
    assert m._halos() == (2 + 12 + 6 + 4 + 2, 8 + 10 + 4 + 6)
    assert m.encoder.supports_row_limit(256) and m.decoder.supports_row_limit(256)
    assert m._skips_pad_rows(True) is False                      # off by default
    m.skip_pad_rows = True
    assert m._skips_pad_rows(True) is True
    assert m._skips_pad_rows(False) is False                      # teacher-forced forward: every row, like the reference
    m.length_buckets = 2
    assert m._skips_pad_rows(True) is False                      # the bucketed path already cuts the rows
    other = _mk("C1").eval()                                      # dense FFN convolutions: no row-limited FFTBlock
    other.skip_pad_rows = True
    assert not other.decoder.supports_row_limit(other.hparams.decoder_hidden)
    with pytest.raises(NotImplementedError):
        other._skips_pad_rows(True)
    wide = _mk("C3").eval()                                       # d = 768, head_dim 384: the wide row-limited block
    wide.skip_pad_rows = True
    assert wide.encoder.supports_row_limit(768) and wide.decoder.supports_row_limit(768)
    assert wide._skips_pad_rows(True) is True
    for layer in wide.decoder.layers:
        layer.wide_flash_attention = False                        # GEMM-decomposed attention: no row limits
    with pytest.raises(NotImplementedError):
        wide._skips_pad_rows(True)
    m.length_buckets, m.skip_pad_rows = 1, True
    m.set_compute_mode("simt")
    with pytest.raises(NotImplementedError):
        m._skips_pad_rows(True)


def test_graph_signature_tracks_parameter_changes():
    """captured bucket graphs bake in weight addresses and packed copies: any parameter change must drop them"""
    m = _mk("TINY_DW").eval()
    m._graphs_validate()
    m._graphs["sentinel"] = {"seen": 1}
    m._graphs_validate()
    assert "sentinel" in m._graphs
    with torch.no_grad():
        m.linear.bias.add_(1.0)
    m._graphs_validate()
    assert "sentinel" not in m._graphs
    m._graphs["sentinel"] = {"seen": 1}
    m.float()                                                     # _apply: parameters may move
    assert "sentinel" not in m.__dict__.get("_graphs", {})
