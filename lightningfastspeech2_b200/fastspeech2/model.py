"""Host-side mirror of ``litfass.fastspeech2.model`` (reference litfass/fastspeech2/model.py).

Same class names, constructor signatures and state_dict keys as the reference, so
reference checkpoints load unchanged and callers (fastspeech2.py, on_load_checkpoint)
need no edits.  The torch.nn sub-modules (Conv1d, Linear, LayerNorm, MultiheadAttention,
Embedding) are used ONLY as parameter containers -- that is what pins the key names,
shapes and default initialisation -- their forward() is never called: every forward here
launches the hand-written sm_100a kernels of liblfs2.so through ``ops``.

Inference only computes forward; dropout layers of the reference are identity in eval
mode and with p=0, which is the parity protocol.  Modules raise if dropout
would be active (train mode with p>0) rather than silently ignoring it.
"""
import math

import numpy as np
import torch
from torch import nn

from .. import ops


def _version_key(*tensors):
    return (ops.WEIGHTS_EPOCH,) + tuple((t.data_ptr(), t._version, t.device) for t in tensors)


class _PackCache:
    """Kernel-friendly weight repacks, rebuilt when a parameter changes
    (keyed on data_ptr/_version, so load_state_dict / optimizer steps invalidate them)."""

    def __init__(self):
        self._key = None
        self._val = None

    def get(self, params, builder):
        key = _version_key(*params)
        if key != self._key:
            with torch.no_grad():
                self._val = builder()
            self._key = key
        return self._val


COMPUTE_MODES = ("fp32", "bf16", "simt")


def _npass(mode):
    """fp32 = bf16 hi/lo split, 3 tensor-core passes per product (fp32 parity, <=1e-3 on mel);
    bf16 = single pass on the hi planes (<=1e-2); simt = exact-fp32 CUDA-core kernels."""
    if mode not in COMPUTE_MODES:
        raise ValueError(f"compute_mode must be one of {COMPUTE_MODES}, got {mode!r}")
    return 3 if mode == "fp32" else 1


def _tc_ok(*dims):
    return all(v % 32 == 0 for v in dims)


def _require_inference(module, p, what):
    if module.training and p > 0:
        raise NotImplementedError(
            f"{what}: dropout p={p} in training mode is not implemented by the CUDA path "
            "(call .eval() or construct with dropout 0)")


class PositionalEncoding(nn.Module):
    """reference model.py:38-55; the add itself is fused into the front-end kernels
    (ops.embed_pe_spk / ops.add_pe_spk_), forward() here serves stand-alone use."""

    def __init__(self, d_model, max_len=5000, dropout=0.1):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0))

    def forward(self, x):
        _require_inference(self, self.dropout.p, "PositionalEncoding")
        zero = torch.zeros(x.shape[0], x.shape[2], device=x.device, dtype=torch.float32)
        return ops.add_pe_spk_(x.contiguous().clone(), self.pe, zero)


class Transpose(nn.Module):
    """reference model.py:58-64 -- kept for state_dict key compatibility (``.module``)."""

    def __init__(self, module):
        super().__init__()
        self.module = module


class ConformerEncoderLayer(nn.Module):
    """FFTBlock, reference model.py:67-122 (post-norm, relu, LayerNorm eps 1e-5).

    Constructor mirrors ``ConformerEncoderLayer(d_model, nhead, conv_in=, conv_filter_size=,
    conv_kernel=(k1,k2), batch_first=True, dropout=, conv_depthwise=)``."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu",
                 layer_norm_eps=1e-5, batch_first=False, norm_first=False, **kwargs):
        super().__init__()
        if not batch_first:
            raise NotImplementedError("batch_first=False")
        if norm_first:
            raise NotImplementedError("norm_first=True (the reference never uses it)")
        if activation not in ("relu", torch.nn.functional.relu):
            raise NotImplementedError("only relu")
        conv_in, fsz = kwargs["conv_in"], kwargs["conv_filter_size"]
        k1, k2 = kwargs["conv_kernel"]
        self.depthwise = bool(kwargs.get("conv_depthwise", False))
        self.nhead = nhead
        self.p_drop = dropout
        self.eps = layer_norm_eps
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=True)
        self.norm1 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        self.norm2 = nn.LayerNorm(d_model, eps=layer_norm_eps)
        if self.depthwise:
            self.conv1 = nn.Sequential(nn.Conv1d(conv_in, conv_in, kernel_size=k1, padding="same", groups=conv_in),
                                       nn.Conv1d(conv_in, fsz, 1))
            self.conv2 = nn.Sequential(nn.Conv1d(fsz, fsz, kernel_size=k2, padding="same", groups=conv_in),
                                       nn.Conv1d(fsz, conv_in, 1))
        else:
            self.conv1 = nn.Conv1d(conv_in, fsz, kernel_size=k1, padding="same")
            self.conv2 = nn.Conv1d(fsz, conv_in, kernel_size=k2, padding="same")
        self._pack = _PackCache()
        self._pack_tc = _PackCache()

    compute_mode = "fp32"
    fused_ffn = True
    # Attention operands in compute mode "fp32": "f16" = the QKV GEMM (3-pass split-bf16, fp32 accumulate) writes q, k, v
    # as ONE fp16 plane and Q.K^T / P.V run as single fp16 tensor-core passes (11 significant bits per operand: mel error
    # ~6e-5 against the 1e-3 budget, tools/precision_emulation.py; a third of the attention MMAs and half of the qkv
    # traffic); "x3" = q, k, v and P as bf16 hi/lo planes, three passes per product (~1.4e-5).
    attention_operands = "f16"
    wide_flash_attention = True   # head_dim 256 / 384: lfs2_attention_tc_wide instead of the GEMM-decomposed attention
    # GEMM sites of compute mode "fp32" (fused d = 256 block) that run the 2-pass recipe: the activation operand as ONE
    # fp16 plane against the bf16 hi/lo weight planes (a.w_hi + a.w_lo, two thirds of the tensor work of the 3-pass
    # product).  "qkv": its results are rounded to fp16 for the attention anyway; "ffn": both GEMMs of the fused FFN.
    # Per-site mel cost in tools/precision_emulation.py (all three: ~1.7e-4 against the 1e-3 budget; the out-projection
    # alone would add 1.6e-4 and stays 3-pass).  () = every GEMM 3-pass.
    two_pass_sites = ("qkv", "ffn")

    # -- weight repacks -------------------------------------------------------------------
    def _build_pack_tc(self):
        """bf16 hi/lo planes of every GEMM weight (operands of lfs2_gemm_tc)."""
        p = self._packed()
        sa = self.self_attn
        w = {"in_proj": ops.split_bf16(sa.in_proj_weight.detach().contiguous()),
             "out_proj": ops.split_bf16(sa.out_proj.weight.detach().contiguous())}
        if self.depthwise:
            w["pw1"] = ops.split_bf16(p["pw1_w"])
            w["w_eff"] = ops.split_bf16(p["w_eff"])
            # fp16 hi/lo planes for the 2-pass recipe (fp16 activation plane; see two_pass_sites)
            w["in_proj16"] = ops.split_f16(sa.in_proj_weight.detach().contiguous())
            w["pw1_16"] = ops.split_f16(p["pw1_w"])
            w["w_eff16"] = ops.split_f16(p["w_eff"])
        else:
            w["c1"] = ops.split_bf16(p["c1_wp"])
            w["c2"] = ops.split_bf16(p["c2_wp"])
        return w

    def _conv_params(self):
        cp = self.__dict__.get("_conv_param_list")
        if cp is None:  # plain attribute (not a registered parameter list): the module tree is walked once
            cp = list(self.conv1.parameters()) + list(self.conv2.parameters())
            self.__dict__["_conv_param_list"] = cp
        return cp

    def _packed_tc(self):
        sa = self.self_attn
        return self._pack_tc.get([sa.in_proj_weight, sa.out_proj.weight] + self._conv_params(), self._build_pack_tc)

    def _build_pack(self):
        p = {}
        if self.depthwise:
            dw, pw = self.conv1[0], self.conv1[1]
            gc, pw2 = self.conv2[0], self.conv2[1]
            if gc.kernel_size[0] != 1:
                raise NotImplementedError("grouped conv2.0 with kernel > 1")
            p["dw_wt"] = dw.weight[:, 0, :].t().contiguous()           # (k, d)
            p["pw1_w"] = pw.weight[:, :, 0].contiguous()               # (F, d)
            # conv2.0 (F->F, groups=d, 1x1) followed by conv2.1 (F->d, 1x1) with nothing in
            # between is one linear map: W_eff = W21 . blockdiag(W20), b_eff = W21.b20 + b21
            # (lfs2_fold_pw_fwd: the same kernel the train step uses; g = F/d terms per element)
            p["w_eff"], p["b_eff"] = ops.fold_pw(pw2.weight[:, :, 0].contiguous(), gc.weight[:, :, 0].contiguous(),
                                                 gc.bias, pw2.bias)
        else:
            fsz, d, k1 = self.conv1.weight.shape
            p["c1_wp"] = self.conv1.weight.permute(0, 2, 1).reshape(fsz, k1 * d).contiguous()
            d2, f2, k2 = self.conv2.weight.shape
            p["c2_wp"] = self.conv2.weight.permute(0, 2, 1).reshape(d2, k2 * f2).contiguous()
        return p

    def _packed(self):
        return self._pack.get(self._conv_params(), self._build_pack)

    def forward(self, src, src_mask=None, src_key_padding_mask=None):
        if src_mask is not None:
            raise NotImplementedError("attention src_mask (the reference never passes one)")
        _require_inference(self, self.p_drop, "ConformerEncoderLayer")
        x = src
        sa = self.self_attn
        if self.tc_capable(x.shape[-1]):
            two = self.compute_mode == "fp32" and "qkv" in self.two_pass_sites and x.shape[-1] == 256
            return ops.merge_planes(self.forward_planes(ops.planes_of(x, want_f16=two), src_key_padding_mask, next_f16=False))
        qkv = ops.linear(x, sa.in_proj_weight, sa.in_proj_bias, tag="qkv_gemm")
        ctx = ops.attention(qkv, src_key_padding_mask, self.nhead)
        a = ops.linear(ctx, sa.out_proj.weight, sa.out_proj.bias, tag="out_proj_gemm")
        x1 = ops.add_layernorm(x, a, self.norm1.weight, self.norm1.bias, self.eps)
        y = self._ff_block(x1)
        return ops.add_layernorm(x1, y, self.norm2.weight, self.norm2.bias, self.eps)

    def halo(self):
        """frames of context on each side that the FFN convolutions of this block reach"""
        if self.depthwise:
            return (self.conv1[0].kernel_size[0] - 1) // 2 + (self.conv2[0].kernel_size[0] - 1) // 2
        return (self.conv1.kernel_size[0] - 1) // 2 + (self.conv2.kernel_size[0] - 1) // 2

    def tc_capable(self, d):
        fsz = self.conv1[1].weight.shape[0] if self.depthwise else self.conv1.weight.shape[0]
        return self.compute_mode != "simt" and _tc_ok(d, fsz)

    def supports_row_limit(self, d):
        """the PAD-row skipping path exists for the fully fused block (d = 256, head_dim 128, depthwise fused FFN) and
        for the wide block (d != 256 with the flash attention for head_dim 256 / 384: the 76 M configuration)"""
        if self.depthwise and self.tc_capable(d) and d != 256:
            return (d // self.nhead) in (256, 384) and self.wide_flash_attention and \
                (self.compute_mode == "bf16" or self.attention_operands == "f16")
        if not (self.depthwise and self.tc_capable(d) and d == 256 and d // self.nhead == 128 and self.fused_ffn):
            return False
        fsz = self.conv1[1].weight.shape[0]
        return fsz % 256 == 0 and fsz <= 2048 and self.conv1[0].kernel_size[0] <= 25

    def forward_planes(self, xp, kpm, row_limit=None, next_f16=True):
        """tcgen05 path, planes in -> planes out.  Activations travel between kernels only as bf16
        hi/lo planes (x = hi + lo to 2^-17): every GEMM has fused bias/ReLU epilogues, the
        residual add rides the tensor core (identity slabs) and LayerNorm is the epilogue of the
        out-proj and FFN-2 GEMMs; attention keeps Q/P in tensor memory.
        row_limit = (lengths int32 (B), extra, cache dict): every kernel of the block skips the 128-row tiles at or
        after lengths[b] + extra of utterance b (see FastSpeech2.skip_pad_rows).
        next_f16: another block follows, so (2-pass QKV recipe) the result also carries its fp16 plane."""
        npass = _npass(self.compute_mode)
        sa, w, p = self.self_attn, self._packed_tc(), self._packed()
        d = xp.shape[-1]
        if row_limit is not None and not self.supports_row_limit(d):
            raise NotImplementedError("row-limited FFTBlock needs d = 256, head_dim 128 and the fused depthwise FFN, or "
                                      "the wide block (head_dim 256 / 384, depthwise FFN, flash attention)")
        if d != 256:
            return self._forward_tc_unfused_ln(xp, kpm, npass, row_limit)
        two = self.two_pass_sites if npass == 3 else ()
        two = two if "pw1_16" in w else ()
        if d // self.nhead == 128:
            f16 = self.compute_mode == "fp32" and self.attention_operands == "f16"
            qkv_pass = 2 if (f16 and "qkv" in two and xp.h is not None) else npass
            qkv = ops.gemm_tc(xp, w["in_proj16" if qkv_pass == 2 else "in_proj"], sa.in_proj_bias,
                              out="f16" if f16 else "planes", npass=qkv_pass, tag="qkv_gemm", row_limit=row_limit)
            _, ctx = ops.attention_tc(qkv, kpm, self.nhead, npass=npass, row_limit=row_limit)
        else:
            ctx = self._attention_any_head_dim(xp, kpm, npass)
        x1p = ops.gemm_tc(ctx, w["out_proj"], sa.out_proj.bias, residual=xp, gamma=self.norm1.weight,
                          beta=self.norm1.bias, eps=self.eps, out="planes", npass=npass, tag="out_proj_ln_gemm",
                          row_limit=row_limit)
        if self.depthwise:
            fsz = w["pw1"].shape[0]
            if self.fused_ffn and fsz % 256 == 0 and fsz <= 2048:
                # FFN-1 -> ReLU -> FFN-2 -> + x1 -> LayerNorm in one kernel, the F-wide intermediate in tensor memory
                ffn_pass = 2 if "ffn" in two else npass
                up = ops.dwconv1d_planes(x1p, p["dw_wt"], self.conv1[0].bias, row_limit=row_limit,
                                         out="f16" if ffn_pass == 2 else "planes")
                w1, w2 = (w["pw1_16"], w["w_eff16"]) if ffn_pass == 2 else (w["pw1"], w["w_eff"])
                return ops.ffn_fused_tc(up, w1, self.conv1[1].bias, w2, p["b_eff"], x1p,
                                        self.norm2.weight, self.norm2.bias, self.eps, npass=ffn_pass, row_limit=row_limit,
                                        want_f16=next_f16 and "qkv" in two)
            up = ops.dwconv1d_planes(x1p, p["dw_wt"], self.conv1[0].bias, row_limit=row_limit)
            vp = ops.gemm_tc(up, w["pw1"], self.conv1[1].bias, relu=True, out="planes", npass=npass, tag="ffn1_gemm")
            w2, b2, taps2 = w["w_eff"], p["b_eff"], 1
        else:
            vp = ops.gemm_tc(x1p, w["c1"], self.conv1.bias, taps=self.conv1.kernel_size[0], relu=True, out="planes",
                             npass=npass, tag="ffn1_gemm")
            w2, b2, taps2 = w["c2"], self.conv2.bias, self.conv2.kernel_size[0]
        return ops.gemm_tc(vp, w2, b2, taps=taps2, residual=x1p, gamma=self.norm2.weight, beta=self.norm2.bias,
                           eps=self.eps, out="planes", npass=npass, tag="ffn2_ln_gemm")

    def _attention_any_head_dim(self, xp, kpm, npass, row_limit=None):
        """head_dim != 128 (e.g. 384 of the 76 M config): attention as tcgen05 GEMMs (Q.K^T, P.V through
        lfs2_gemm_tc2) around a row-softmax kernel; head_dim not a multiple of 32: CUDA-core flash kernel.
        Returns ctx as Planes."""
        sa, w = self.self_attn, self._packed_tc()
        d = xp.shape[-1]
        if (d // self.nhead) in (256, 384) and self.wide_flash_attention:
            # wide heads (76 M configuration: head_dim 384): flash attention on ONE 16-bit plane of q, k, v -- fp16 in
            # compute mode "fp32" (attention_operands "f16"), bf16 in "bf16" mode; no T x T tensor in HBM
            f16 = self.compute_mode == "fp32"
            if not f16 or self.attention_operands == "f16":
                qkv = ops.gemm_tc(xp, w["in_proj"], sa.in_proj_bias, out="f16" if f16 else "bf16", npass=npass,
                                  tag="qkv_gemm", row_limit=row_limit)
                return ops.attention_tc_wide(qkv, kpm, self.nhead, row_limit=row_limit)[1]
        if row_limit is not None:
            raise NotImplementedError("row limits need the flash attention kernels")
        if (d // self.nhead) % 32 == 0:
            qkv = ops.gemm_tc(xp, w["in_proj"], sa.in_proj_bias, out="planes", npass=npass, tag="qkv_gemm")
            ctx, _, _ = ops.attention_mat_fwd(qkv, kpm, self.nhead, npass=npass)
            return ops.split_bf16(ctx)
        qkv = ops.gemm_tc(xp, w["in_proj"], sa.in_proj_bias, npass=npass, tag="qkv_gemm")
        return ops.split_bf16(ops.attention(qkv, kpm, self.nhead))

    def _forward_tc_unfused_ln(self, xp, kpm, npass, row_limit=None):
        """d != 256 (e.g. the 76 M config, d = 768): tensor-core GEMMs with fp32 results, LayerNorm as its own kernel
        that takes the residual stream as planes and hands planes back -- the block is planes in, planes out.
        In "bf16" mode results that only feed single-pass products (the FFN intermediate) travel as ONE bf16 plane."""
        sa, w, p = self.self_attn, self._packed_tc(), self._packed()
        ctx = self._attention_any_head_dim(xp, kpm, npass, row_limit)
        a = ops.gemm_tc(ctx, w["out_proj"], sa.out_proj.bias, npass=npass, tag="out_proj_gemm", row_limit=row_limit,
                        zero_skipped=False)
        x1p = ops.add_layernorm_planes(xp, a, self.norm1.weight, self.norm1.bias, self.eps, row_limit=row_limit)
        one = "bf16" if npass == 1 else "planes"
        if self.depthwise:
            up = ops.dwconv1d_planes(x1p, p["dw_wt"], self.conv1[0].bias, row_limit=row_limit)
            vp = ops.gemm_tc(up, w["pw1"], self.conv1[1].bias, relu=True, out=one, npass=npass, tag="ffn1_gemm",
                             row_limit=row_limit)
            y = ops.gemm_tc(vp, w["w_eff"], p["b_eff"], npass=npass, tag="ffn2_gemm", row_limit=row_limit,
                            zero_skipped=False)
        elif row_limit is not None:
            raise NotImplementedError("row limits need the depthwise FFN")
        else:
            vp = ops.gemm_tc(x1p, w["c1"], self.conv1.bias, taps=self.conv1.kernel_size[0], relu=True, out=one,
                             npass=npass, tag="ffn1_gemm")
            y = ops.gemm_tc(vp, w["c2"], self.conv2.bias, taps=self.conv2.kernel_size[0], npass=npass, tag="ffn2_gemm")
        return ops.add_layernorm_planes(x1p, y, self.norm2.weight, self.norm2.bias, self.eps, row_limit=row_limit)

    def _ff_block(self, x):
        p = self._packed()
        if self.depthwise:
            u = ops.dwconv1d(x, p["dw_wt"], self.conv1[0].bias)
            v = ops.linear(u, p["pw1_w"], self.conv1[1].bias, relu=True, tag="ffn1_gemm")
            return ops.linear(v, p["w_eff"], p["b_eff"], tag="ffn2_gemm")
        v = ops.conv1d_dense(x, p["c1_wp"], self.conv1.bias, self.conv1.kernel_size[0], relu=True)
        return ops.conv1d_dense(v, p["c2_wp"], self.conv2.bias, self.conv2.kernel_size[0])


class SpeakerEmbedding(nn.Module):
    """reference model.py:125-143.  Only d-vector speakers work at the reference HEAD
    (SURVEY 8 quirk 2); forward returns the (B, d) term, broadcast over time is fused
    into the consumer kernels when called through FastSpeech2.forward."""

    def __init__(self, embedding_dim, speaker_type, nspeakers=None):
        super().__init__()
        self.speaker_type = speaker_type
        self.embedding_dim = embedding_dim
        if "dvector" in speaker_type:
            self.projection = nn.Linear(256, embedding_dim)
            self.has_projection = True
        elif speaker_type == "id":
            self.speaker_embedding = nn.Embedding(nspeakers, embedding_dim)
        self.relu = nn.ReLU()

    def project(self, x):
        if not getattr(self, "has_projection", False):
            raise NotImplementedError("speaker_type='id' (raises AttributeError in the reference too)")
        return ops.speaker_proj(x.contiguous(), self.projection.weight, self.projection.bias)

    def forward(self, x, input_length, output_shape):
        out = self.project(x)
        return out.reshape(-1, 1, output_shape).expand(-1, input_length, -1)


class PriorEmbedding(nn.Module):
    """reference model.py:146-164: per-utterance scalar prior -> bucketize -> embedding -> relu, broadcast over time."""

    def __init__(self, embedding_dim, nbins, stats):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.bins = nn.Parameter(torch.linspace(stats["min"], stats["max"], nbins - 1), requires_grad=False)
        self.embedding = nn.Embedding(nbins, embedding_dim)
        self.relu = nn.ReLU()

    def term(self, x):
        """(B) prior values -> ((B, d) term, (B) bucket indices)"""
        x = torch.as_tensor(x, dtype=torch.float32).to(self.bins.device).contiguous()
        return ops.prior_embed(x, self.bins, self.embedding.weight)

    def forward(self, x, input_length):
        out, _ = self.term(x)
        return out.reshape(-1, 1, self.embedding_dim).expand(-1, input_length, -1)


class VarianceConvolutionLayer(nn.Module):
    """reference model.py:524-561: Transpose(conv) -> ReLU -> LayerNorm(filter) -> Dropout."""

    def __init__(self, in_channels, filter_size, kernel_size, dropout, depthwise):
        super().__init__()
        self.depthwise = depthwise
        self.kernel_size = kernel_size
        if not depthwise:
            conv = nn.Conv1d(in_channels, filter_size, kernel_size, padding=(kernel_size - 1) // 2)
        else:
            conv = nn.Sequential(
                nn.Conv1d(in_channels, in_channels, kernel_size, padding=(kernel_size - 1) // 2, groups=in_channels),
                nn.Conv1d(in_channels, filter_size, 1))
        self.layers = nn.Sequential(Transpose(conv), nn.ReLU(), nn.LayerNorm(filter_size), nn.Dropout(dropout))
        self._pack = _PackCache()

    compute_mode = "fp32"

    def _build_pack(self):
        conv = self.layers[0].module
        if self.depthwise:
            p = {"dw_wt": conv[0].weight[:, 0, :].t().contiguous(), "pw_w": conv[1].weight[:, :, 0].contiguous()}
            if conv[1].weight.is_cuda:
                p["pw_planes"] = ops.split_bf16(p["pw_w"])
            return p
        f, d, k = conv.weight.shape
        p = {"wp": conv.weight.permute(0, 2, 1).reshape(f, k * d).contiguous()}
        if conv.weight.is_cuda:
            p["wp_planes"] = ops.split_bf16(p["wp"])
        return p

    def forward(self, x, out="f32", row_limit=None):
        """x: fp32 (B,T,d) or Planes; out="planes" keeps the result as hi/lo planes for the next layer
        (tensor-core path with the fused ReLU+LayerNorm epilogue only).  row_limit: see VariancePredictor."""
        _require_inference(self, self.layers[3].p, "VarianceConvolutionLayer")
        conv, ln = self.layers[0].module, self.layers[2]
        cp = self.__dict__.get("_conv_param_list")
        if cp is None:
            cp = list(conv.parameters())
            self.__dict__["_conv_param_list"] = cp
        p = self._pack.get(cp, self._build_pack)
        fsz = ln.weight.shape[0]
        if self.compute_mode != "simt" and _tc_ok(x.shape[-1], fsz):
            npass = _npass(self.compute_mode)
            fuse = dict(gamma=ln.weight, beta=ln.bias, eps=ln.eps) if fsz == 256 else {}
            out = out if fuse else "f32"
            if self.depthwise:
                lim = row_limit if fuse else None  # (the unfused LayerNorm kernel has no row limit)
                up = ops.dwconv1d_planes(x if isinstance(x, ops.Planes) else x.contiguous(), p["dw_wt"], conv[0].bias,
                                         row_limit=lim)
                h = ops.gemm_tc(up, p["pw_planes"], conv[1].bias, relu=True, npass=npass, out=out,
                                tag="predictor_pw_ln_gemm", row_limit=lim, **fuse)
            else:
                xp = x if isinstance(x, ops.Planes) else ops.planes_of(x)
                h = ops.gemm_tc(xp, p["wp_planes"], conv.bias, taps=self.kernel_size, relu=True, npass=npass, out=out,
                                tag="predictor_conv_ln_gemm", **fuse)
            return h if fuse else ops.add_layernorm(h, None, ln.weight, ln.bias, ln.eps)
        if isinstance(x, ops.Planes):
            x = ops.merge_planes(x)
        if self.depthwise:
            u = ops.dwconv1d(x, p["dw_wt"], conv[0].bias)
            h = ops.linear(u, p["pw_w"], conv[1].bias, relu=True, tag="predictor_pw_gemm")
        else:
            h = ops.conv1d_dense(x, p["wp"], conv.bias, self.kernel_size, relu=True)
        return ops.add_layernorm(h, None, ln.weight, ln.bias, ln.eps)


# ---- stochastic duration predictor, inference direction ---------------------------------------------------
class _LayerNorm2(nn.Module):
    """parameter container of third_party/stochastic_duration_predictor/normalization.py:5-28 (keys gamma / beta)"""

    def __init__(self, channels, eps=1e-5):
        super().__init__()
        self.eps = eps
        self.gamma = nn.Parameter(torch.ones(channels))
        self.beta = nn.Parameter(torch.zeros(channels))


class DilatedDepthSeparableConv(nn.Module):
    """sdp.py:11-70: num_layers x [depthwise conv (dilation k^i) -> LN -> GELU -> 1x1 conv -> LN -> GELU -> + x]"""

    def __init__(self, channels, kernel_size, num_layers, dropout_p=0.0):
        super().__init__()
        self.num_layers, self.p_drop = num_layers, dropout_p
        self.convs_sep, self.convs_1x1 = nn.ModuleList(), nn.ModuleList()
        self.norms_1, self.norms_2 = nn.ModuleList(), nn.ModuleList()
        for i in range(num_layers):
            dil = kernel_size ** i
            self.convs_sep.append(nn.Conv1d(channels, channels, kernel_size, groups=channels, dilation=dil,
                                            padding=(kernel_size * dil - dil) // 2))
            self.convs_1x1.append(nn.Conv1d(channels, channels, 1))
            self.norms_1.append(_LayerNorm2(channels))
            self.norms_2.append(_LayerNorm2(channels))
        self._pack = _PackCache()

    def _build_pack(self):
        return [(c.weight[:, 0, :].t().contiguous(), p.weight[:, :, 0].contiguous()) for c, p in
                zip(self.convs_sep, self.convs_1x1)]

    def forward(self, x, pad_mask, g=None):
        """x (B,T,C) fp32 channels-last [, already x + g: the caller adds the conditioning]; PAD rows are read as zeros by
        every depthwise conv; the reference's trailing `* x_mask` only touches PAD rows, which nothing reads afterwards"""
        if g is not None:
            raise NotImplementedError("pass x + g (lfs2_sdp_flow_pre adds the conditioning)")
        _require_inference(self, self.p_drop, "DilatedDepthSeparableConv")
        packs = self._pack.get([c.weight for c in self.convs_sep] + [c.weight for c in self.convs_1x1], self._build_pack)
        for i in range(self.num_layers):
            sep, pw = self.convs_sep[i], self.convs_1x1[i]
            y = ops.sdp_dwconv(x, pad_mask, packs[i][0], sep.bias, sep.dilation[0])
            y = ops.sdp_ln_gelu(y, self.norms_1[i].gamma, self.norms_1[i].beta, self.norms_1[i].eps)
            y = ops.linear(y, packs[i][1], pw.bias, tag="sdp_1x1")
            x = ops.sdp_ln_gelu(y, self.norms_2[i].gamma, self.norms_2[i].beta, self.norms_2[i].eps, res=x)
        return x


class ElementwiseAffine(nn.Module):
    """sdp.py:73-95 (parameter container; the reverse direction is lfs2_sdp_affine_reverse)"""

    def __init__(self, channels):
        super().__init__()
        self.translation = nn.Parameter(torch.zeros(channels, 1))
        self.log_scale = nn.Parameter(torch.zeros(channels, 1))


class ConvFlow(nn.Module):
    """sdp.py:98-164: spline coupling flow; reverse direction only"""

    def __init__(self, in_channels, hidden_channels, kernel_size, num_layers, num_bins=10, tail_bound=5.0):
        super().__init__()
        if num_bins != 10 or in_channels != 2:
            raise NotImplementedError("the spline kernel is built for 2 channels and 10 bins (the reference's only use)")
        self.num_bins, self.tail_bound, self.hidden_channels = num_bins, tail_bound, hidden_channels
        self.half_channels = in_channels // 2
        self.pre = nn.Conv1d(self.half_channels, hidden_channels, 1)
        self.convs = DilatedDepthSeparableConv(hidden_channels, kernel_size, num_layers, dropout_p=0.0)
        self.proj = nn.Conv1d(hidden_channels, self.half_channels * (num_bins * 3 - 1), 1)
        self.proj.weight.data.zero_()
        self.proj.bias.data.zero_()
        self._pack = _PackCache()

    def _build_pack(self):
        n = self.proj.weight.shape[0]                 # 29 spline parameters, padded to 32 output columns for lfs2_linear
        w = torch.zeros(32, self.hidden_channels, device=self.proj.weight.device)
        b = torch.zeros(32, device=self.proj.weight.device)
        w[:n] = self.proj.weight[:, :, 0]
        b[:n] = self.proj.bias
        return {"pre_w": self.pre.weight[:, 0, 0].contiguous(), "proj_w": w, "proj_b": b}

    def reverse_(self, z, flip, pad_mask, g):
        """in place on the flow state z (B,T,2): logical channel 0 (physical `flip`) conditions, logical 1 is transformed"""
        pk = self._pack.get([self.pre.weight, self.proj.weight, self.proj.bias], self._build_pack)
        h = ops.sdp_flow_pre(z, flip, pk["pre_w"], self.pre.bias, g)
        h = self.convs(h, pad_mask)
        h = ops.linear(h, pk["proj_w"], pk["proj_b"], tag="sdp_proj")
        ops.sdp_spline_inverse_(z, 1 ^ flip, h, pad_mask, self.hidden_channels, self.tail_bound)


class StochasticDurationPredictor(nn.Module):
    """sdp.py:167-349.  All parameters of the reference module exist (a reference checkpoint loads strictly); the CUDA path
    implements the INFERENCE direction (reverse=True).  The training direction (variational dequantisation + the flows'
    negative log-likelihood, sdp.py:271-328) raises NotImplementedError."""

    def __init__(self, in_channels, hidden_channels, kernel_size, dropout_p, num_flows=4, cond_channels=0, language_emb_dim=0):
        super().__init__()
        if cond_channels or language_emb_dim:
            raise NotImplementedError("conditioning / language embeddings (the reference's wrapper never passes them)")
        self.hidden_channels = hidden_channels
        self.pre = nn.Conv1d(in_channels, hidden_channels, 1)
        self.convs = DilatedDepthSeparableConv(hidden_channels, kernel_size, num_layers=3, dropout_p=dropout_p)
        self.proj = nn.Conv1d(hidden_channels, hidden_channels, 1)
        self.flows = nn.ModuleList([ElementwiseAffine(2)] +
                                   [ConvFlow(2, hidden_channels, kernel_size, num_layers=3) for _ in range(num_flows)])
        self.post_pre = nn.Conv1d(1, hidden_channels, 1)
        self.post_convs = DilatedDepthSeparableConv(hidden_channels, kernel_size, num_layers=3, dropout_p=dropout_p)
        self.post_proj = nn.Conv1d(hidden_channels, hidden_channels, 1)
        self.post_flows = nn.ModuleList([ElementwiseAffine(2)] +
                                        [ConvFlow(2, hidden_channels, kernel_size, num_layers=3) for _ in range(num_flows)])
        self._pack = _PackCache()

    def forward(self, x, x_mask, dr=None, g=None, lang_emb=None, reverse=False, noise_scale=1.0, noise=None):
        """x (B,T,C), x_mask (B,T) bool True = PAD -> log-durations (B,T).  noise (B,2,T): the torch.randn draw of
        sdp.py:331 (injected for parity runs; drawn on the device otherwise)."""
        if not reverse or dr is not None:
            raise NotImplementedError("stochastic duration predictor: only the inference direction runs on the CUDA path")
        if g is not None or lang_emb is not None:
            raise NotImplementedError("conditioning inputs")
        pk = self._pack.get([self.pre.weight, self.proj.weight],
                            lambda: (self.pre.weight[:, :, 0].contiguous(), self.proj.weight[:, :, 0].contiguous()))
        b, t, _ = x.shape
        c = ops.linear(x.contiguous(), pk[0], self.pre.bias, tag="sdp_pre")
        c = self.convs(c, x_mask)
        c = ops.linear(c, pk[1], self.proj.bias, tag="sdp_proj")
        if noise is None:
            noise = torch.randn(b, 2, t, device=x.device, dtype=torch.float32)
        z = (noise.to(x.device, torch.float32) * noise_scale).transpose(1, 2).contiguous()   # (B,T,2)
        order = list(range(len(self.flows)))[::-1]
        order = order[:-2] + [order[-1]]          # sdp.py:331: "remove a useless vflow"
        flip = 0
        for j in order:
            flip ^= 1                              # torch.flip(z, [1]) before every flow: tracked, not materialised
            if j == 0:
                ea = self.flows[0]
                ops.sdp_affine_reverse_(z, ea.translation, ea.log_scale, x_mask, flip)
            else:
                self.flows[j].reverse_(z, flip, x_mask, c)
        return z[:, :, flip].contiguous()         # logical channel 0


class StochasticDurationPredictorWrapper(nn.Module):
    """reference model.py:463-480 (same constructor; `nlayers` is the number of flows, as in the reference's call)."""

    def __init__(self, nlayers, in_channels, filter_size, kernel_size, dropout):
        super().__init__()
        self.sdp = StochasticDurationPredictor(in_channels, filter_size, kernel_size, dropout, nlayers)

    def forward(self, x, mask, tgt=None, sigma=1.0, inference=False, noise=None):
        if isinstance(x, ops.Planes):
            x = ops.merge_planes(x)
        out = self.sdp(x, mask, tgt, reverse=inference, noise_scale=sigma, noise=noise)
        if mask is not None and inference:
            pass                                    # PAD rows are already zeros: every flow step writes them as zeros
        return out


class VariancePredictor(nn.Module):
    """reference model.py:482-522."""

    def halo(self):
        """rows of context on each side that one output row depends on"""
        return sum((layer.kernel_size - 1) // 2 for layer in self.layers)

    def __init__(self, nlayers, in_channels, filter_size, kernel_size, dropout, depthwise=False, cwt=False):
        super().__init__()
        if cwt:
            raise NotImplementedError("cwt variance transform (needs scipy.signal.cwt, removed upstream)")
        self.layers = nn.Sequential(*[
            VarianceConvolutionLayer(in_channels, filter_size, kernel_size, dropout, depthwise)
            for _ in range(nlayers)])
        self.cwt = cwt
        self.linear = nn.Linear(filter_size, 1)

    skip_pad_tiles = True
    fused_layers = True   # depthwise k = 3 stacks at width 256: each layer's GEMM epilogue also does the NEXT layer's conv / the head

    def _fusable(self, x):
        if not (self.fused_layers and not isinstance(x, ops.Planes) and x.dim() == 3 and x.shape[-1] == 256):
            return False
        for layer in self.layers:
            ln = layer.layers[2]
            if not (layer.depthwise and layer.kernel_size == 3 and layer.compute_mode != "simt"
                    and ln.weight.shape[0] == 256 and not (layer.training and layer.layers[3].p > 0)):
                return False
        return self.linear.weight.shape[1] == 256

    def _forward_fused(self, x, mask, lim):
        """dwconv(layer 0) as its own kernel, then ONE launch per layer: GEMM + bias + ReLU + LayerNorm + the next layer's
        depthwise conv (neighbour rows from the neighbouring epilogue threads) -- and, in the last layer, the Linear(256, 1)
        head + mask instead.  Per layer the activations cross HBM once each way instead of twice."""
        layers = list(self.layers)
        packs = []
        for layer in layers:
            conv = layer.layers[0].module
            cp = layer.__dict__.get("_conv_param_list")
            if cp is None:
                cp = list(conv.parameters())
                layer.__dict__["_conv_param_list"] = cp
            packs.append(layer._pack.get(cp, layer._build_pack))
        npass = _npass(layers[0].compute_mode)
        up = ops.dwconv1d_planes(x.contiguous(), packs[0]["dw_wt"], layers[0].layers[0].module[0].bias, row_limit=lim)
        for i, layer in enumerate(layers):
            conv, ln = layer.layers[0].module, layer.layers[2]
            if i + 1 < len(layers):
                nxt = layers[i + 1].layers[0].module
                up = ops.predictor_layer_tc(up, packs[i]["pw_planes"], conv[1].bias, ln.weight, ln.bias, ln.eps, npass,
                                            next_dw=(packs[i + 1]["dw_wt"], nxt[0].bias), row_limit=lim)
            else:
                return ops.predictor_layer_tc(up, packs[i]["pw_planes"], conv[1].bias, ln.weight, ln.bias, ln.eps, npass,
                                              head=(self.linear.weight, self.linear.bias, mask), row_limit=lim)

    def forward(self, x, mask=None, return_conv=False):
        """The head masks every PAD position to 0 (model.py:518), so only rows within the conv halo of an utterance's
        last valid position can influence the result: 128-row tiles beyond that are skipped in the depthwise /
        tensor-core path (bit-identical output; the hidden state `z` of skipped rows is undefined)."""
        z = x
        nl = len(self.layers)
        lim = None
        if (self.skip_pad_tiles and mask is not None and not return_conv and not isinstance(x, ops.Planes)
                and x.dim() == 3 and all(l.depthwise and l.compute_mode != "simt" for l in self.layers)):
            lim = (ops.mask_lengths(mask), self.halo(), {})  # {} = tile list shared by the layers of this call
        if not return_conv and self._fusable(x):
            return self._forward_fused(x, mask, lim)
        for i, layer in enumerate(self.layers):
            z = layer(z, out="planes" if i + 1 < nl else "f32", row_limit=lim)
        if isinstance(z, ops.Planes):
            z = ops.merge_planes(z)
        out = ops.rowdot_mask(z, self.linear.weight, self.linear.bias, mask)
        return (out, z) if return_conv else out


class VarianceEncoder(nn.Module):
    """reference model.py:373-461 (non-CWT branch)."""

    def __init__(self, nlayers, in_channels, filter_size, kernel_size, dropout, depthwise, min, max, mean, std,
                 nbins, cwt):
        super().__init__()
        if cwt:
            raise NotImplementedError("cwt variance transform")
        self.cwt = cwt
        self.predictor = VariancePredictor(nlayers, in_channels, filter_size, kernel_size, dropout, depthwise, cwt)
        self.bins = nn.Parameter(torch.linspace(min, max, nbins - 1), requires_grad=False)
        self.embedding = nn.Embedding(nbins, in_channels)
        self.mean = mean
        self.std = std

    def encode_(self, x, tgt, mask, control=1.0, acc=None, acc_init=False, forced_idx=None, want_idx=False, tail=None):
        """Fused form used by VarianceAdaptor: predicts, bucketizes and does ``x += emb``
        in place (and ``acc (+)= emb``).  Returns (prediction, bucket indices or None).
        tail = (pe, spk, want_f16): this is the last encoder before the decoder -- x is left untouched and
        ((x + emb) + pe) + spk is returned as a third value, the decoder's input Planes (ops.decoder_input_planes)."""
        prediction = self.predictor(x, mask)
        val = prediction if tgt is None else tgt.to(device=x.device, dtype=torch.float32).contiguous()
        if tail is not None:
            planes, idx = ops.decoder_input_planes(
                x, tail[0], tail[1], want_f16=tail[2],
                bucket=dict(val=val, std=self.std, mean=self.mean, bins=self.bins, emb=self.embedding.weight,
                            idx_forced=forced_idx, acc=acc, acc_init=acc_init, want_idx=want_idx))
            if tgt is None and control != 1.0:
                prediction = prediction * control
            return prediction, idx, planes
        idx = ops.bucket_embed_add_(x, val, self.std, self.mean, self.bins, self.embedding.weight,
                                    idx_forced=forced_idx, acc=acc, acc_init=acc_init, want_idx=want_idx)
        if tgt is None and control != 1.0:
            prediction = prediction * control
        return prediction, idx

    def forward(self, x, tgt, mask, control=1.0):
        """Reference signature: returns (prediction, embedding)."""
        # stand-alone use: compute the embedding into a zero tensor without touching x
        emb = torch.zeros_like(x)
        prediction = self.predictor(x, mask)
        val = prediction if tgt is None else tgt.to(device=x.device, dtype=torch.float32).contiguous()
        ops.bucket_embed_add_(emb, val, self.std, self.mean, self.bins, self.embedding.weight)
        if tgt is None:
            prediction = prediction * control
        return prediction, emb


class LengthRegulator(nn.Module):
    """reference model.py:344-370."""

    def __init__(self, pad_to_multiple_of=None):
        super().__init__()
        self.pad_to_multiple_of = pad_to_multiple_of

    def forward(self, x, durations, max_length=None, scan=None, frames=None):
        """pad_to_multiple_of (model.py:356-357, 362-366; the FastDiff adaptor's regulator): the output length
        min(longest, int(max_length)) is rounded UP to the multiple, and frames up to that rounded length are kept --
        an utterance longer than int(max_length) keeps its frames up to the rounded length, like the reference's
        pad_sequence + cut."""
        return ops.length_regulate(x.contiguous(), durations.to(x.device), max_length, scan=scan, frames=frames,
                                   pad_to_multiple_of=self.pad_to_multiple_of)


class VarianceAdaptor(nn.Module):
    """reference model.py:167-341."""

    def __init__(self, stats, variances, variance_levels, variance_transforms, variance_nlayers,
                 variance_kernel_size, variance_dropout, variance_filter_size, variance_nbins,
                 variance_depthwise_conv, duration_nlayers, duration_stochastic, duration_kernel_size,
                 duration_dropout, duration_filter_size, duration_depthwise_conv, encoder_hidden, max_length):
        super().__init__()
        self.variances = variances
        self.variance_levels = variance_levels
        self.variance_transforms = variance_transforms
        self.duration_stochastic = duration_stochastic
        self.max_length = max_length
        if duration_stochastic:
            if duration_depthwise_conv:   # reference model.py:197-200
                raise NotImplementedError("Depthwise convolution not implemented for Flow-Based duration prediction")
            self.duration_predictor = StochasticDurationPredictorWrapper(duration_nlayers, encoder_hidden,
                                                                         duration_filter_size, duration_kernel_size,
                                                                         duration_dropout)
        else:
            self.duration_predictor = VariancePredictor(duration_nlayers, encoder_hidden, duration_filter_size,
                                                        duration_kernel_size, duration_dropout, duration_depthwise_conv)
        self.length_regulator = LengthRegulator()
        encoders = {}
        for i, var in enumerate(variances):
            encoders[var] = VarianceEncoder(variance_nlayers[i], encoder_hidden, variance_filter_size,
                                            variance_kernel_size[i], variance_dropout[i], variance_depthwise_conv,
                                            stats[var]["min"], stats[var]["max"], stats[var]["mean"],
                                            stats[var]["std"], variance_nbins,
                                            cwt=variance_transforms[i] == "cwt")
        self.encoders = nn.ModuleDict(encoders)
        self.frozen_components = []

    @staticmethod
    def _fit_forced(idx, width, enc):
        """forced bucket indices cut / extended to `width` frames; frames past the given ones are PAD frames,
        whose prediction is masked to 0, i.e. bucket(mean) (parity harness only)"""
        if idx.shape[1] >= width:
            return idx[:, :width].contiguous()
        pad = int(torch.bucketize(torch.tensor(0.0 * enc.std + enc.mean), enc.bins.detach().cpu()))
        ext = torch.full((idx.shape[0], width - idx.shape[1]), pad, device=idx.device, dtype=idx.dtype)
        return torch.cat([idx, ext], 1).contiguous()

    # result["out"] (sum of the variance embeddings) only feeds the optional fastdiff head (fastspeech2.py:733-736);
    # FastSpeech2 clears this when the head does not exist
    need_out_val = True

    def freeze(self, component):
        mod = self.duration_predictor if component == "duration" else self.encoders[component]
        for param in mod.parameters():
            param.requires_grad = False
        self.frozen_components.append(component)

    def forward(self, x, src_mask, targets, inference=False, tf_ratio=1.0, oracles=[], force=None, control=None):
        st = self.durations(x, src_mask, targets, inference=inference, tf_ratio=tf_ratio, force=force, control=control,
                            oracles=oracles)
        return self.expand(st["x_phone"], st, targets, inference=inference, oracles=oracles, force=force,
                           control=control)

    def durations(self, x, src_mask, targets, inference=False, tf_ratio=1.0, force=None, control=None, oracles=[]):
        """first half of the reference forward (model.py:249-309): duration prediction and the durations used"""
        force = force or {}
        control = control or {}
        if not self.duration_stochastic:
            duration_pred = self.duration_predictor(x, src_mask)
        elif inference:   # model.py:265-268 (x.detach(): no gradient path exists here anyway)
            duration_pred = self.duration_predictor(x, src_mask, inference=True, noise=force.get("sdp_noise"),
                                                    sigma=force.get("sdp_sigma", 1.0))
        else:
            raise NotImplementedError("stochastic duration predictor: the training direction (negative log-likelihood of "
                                      "the flows, sdp.py:271-328) is not implemented on the CUDA path")
        tf_val = np.random.uniform(0, 1) <= tf_ratio  # reference model.py:272
        # phone-level variances act on the encoder output before the LengthRegulator (model.py:277-294)
        result, out_val = {}, None
        phone_vars = [(i, v) for i, v in enumerate(self.variances) if self.variance_levels[i] == "phone"]
        if phone_vars:
            x = x.clone() if not isinstance(x, ops.Planes) else ops.merge_planes(x)  # duration_pred read the original
            out_val = torch.empty_like(x)
            for n, (i, var) in enumerate(phone_vars):
                teacher = (not inference and tf_val) or var in oracles
                forced = force.get("bucket_idx", {}).get(var)
                pred, idx = self.encoders[var].encode_(
                    x, targets[f"variances_{var}"] if teacher else None, src_mask, control.get(var, 1.0), acc=out_val,
                    acc_init=(n == 0), forced_idx=None if forced is None else forced.to(x.device).contiguous(),
                    want_idx=force.get("want_idx", False))
                result[f"variances_{var}"] = pred
                if idx is not None:
                    result[f"_bucket_{var}"] = idx
        if "duration_rounded" in force:
            duration_rounded = force["duration_rounded"].to(x.device)
        elif not inference:
            duration_rounded = targets["duration"].to(x.device)
        elif self.duration_stochastic:
            duration_rounded = ops.sdp_durations(duration_pred, src_mask)
        else:
            duration_rounded = ops.duration_round_guard(duration_pred, src_mask)
        return {"duration_prediction": duration_pred, "duration_rounded": duration_rounded, "tf_val": tf_val,
                "x_phone": x, "out_phone": out_val, "phone_result": result}

    def expand(self, x, st, targets, inference=False, oracles=[], force=None, control=None, scan=None, frames=None,
               tail=None):
        """second half (model.py:311-341): LengthRegulator, then the frame-level variance encoders in sequence.
        tail = (pe, spk, want_f16): the caller only needs the decoder's input ((x + pe) + spk) as Planes -- the last
        embedding add, the positional / speaker add and the plane split run as one kernel; result["x_planes"]."""
        force = force or {}
        control = control or {}
        result = dict(st.get("phone_result") or {})
        duration_pred, duration_rounded, tf_val = st["duration_prediction"], st["duration_rounded"], st["tf_val"]
        if scan is None:
            scan = ops.length_regulate_scan(duration_rounded.to(x.device), x.shape[:2])
        lens_host = None
        if frames is None:
            # the single device->host sync of the path: the per-utterance frame counts (B int64; their maximum sizes the
            # frame-level tensors, the list itself lets a caller read back only the valid frames, pipeline.SynthesisStream)
            lens_host = scan[1].tolist()
            longest = max(lens_host) if lens_host else 0
            l = min(longest, int(self.max_length)) if self.max_length is not None else longest
            frames = (l, l)
        x_in = x
        x, tgt_mask = self.length_regulator(x_in, duration_rounded, self.max_length, scan=scan, frames=frames)
        have_acc = st.get("out_phone") is not None
        if not self.need_out_val:
            have_acc, out_val = False, None  # nobody reads result["out"] (no fastdiff head): skip its HBM traffic
        elif have_acc:  # the summed phone-level embeddings are length-regulated too (model.py:312-313)
            out_val, _ = self.length_regulator(st["out_phone"], duration_rounded, self.max_length, scan=scan, frames=frames)
        else:
            out_val = torch.empty_like(x) if len(self.variances) else None
        nframe = 0
        frame_vars = [i for i in range(len(self.variances)) if self.variance_levels[i] == "frame"]
        x_planes = None
        for i, var in enumerate(self.variances):
            if self.variance_levels[i] != "frame":
                continue
            teacher = (not inference and tf_val) or var in oracles
            tgt = targets[f"variances_{var}"] if teacher else None
            forced = force.get("bucket_idx", {}).get(var)
            if forced is not None:
                forced = self._fit_forced(forced.to(x.device), x.shape[1], self.encoders[var])
            enc = self.encoders[var].encode_(
                x, tgt, tgt_mask, control.get(var, 1.0), acc=out_val, acc_init=(nframe == 0 and not have_acc),
                forced_idx=forced,
                want_idx=force.get("want_idx", False), tail=tail if i == frame_vars[-1] else None)
            pred, idx = enc[0], enc[1]
            if len(enc) > 2:
                x_planes = enc[2]
            nframe += 1
            result[f"variances_{var}"] = pred
            if idx is not None:
                result[f"_bucket_{var}"] = idx

        if tail is not None and x_planes is None:  # no frame-level variance: positional / speaker add + split only
            x_planes, _ = ops.decoder_input_planes(x, tail[0], tail[1], want_f16=tail[2])
        result["x"] = x
        result["x_planes"] = x_planes
        result["frame_lengths"] = scan[1]            # (B) int64 on the device: frames per utterance before the max_length cut
        result["frame_lengths_host"] = lens_host     # the same as a Python list when this call did the sync, else None
        result["duration_prediction"] = duration_pred
        result["duration_rounded"] = duration_rounded
        result["tgt_mask"] = tgt_mask
        result["out"] = out_val
        return result
