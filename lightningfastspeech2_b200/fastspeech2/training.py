"""Train-step graph of the mel-generation path: forward that saves what the backward needs,
and the hand-written backward, both as sequences of liblfs2.so kernels (SURVEY.md 8a
"Backward notes").  The reference gets all of this from torch.autograd over its nn.Modules
(litfass/fastspeech2/fastspeech2.py:786-797 -> forward :636-784); here the graph is fixed and
known, so the backward is written out explicitly: no autograd tape of ATen kernels, every
gradient accumulation is a kernel of this library, parameter gradients accumulate (+=) straight
into ``p.grad`` (views of one flat buffer when ``FastSpeech2.flatten_parameters()`` was called,
which is what the fused AdamW and the single NCCL all-reduce consume).

Dropout: all seven sites of the reference (PE dropout x2, attention probabilities, dropout1/2 and the
post-ReLU dropout of the FFTBlock, predictor layers) are implemented with a counter-based Philox mask
that the backward regenerates from (seed, site); the random stream necessarily differs from PyTorch's,
so gradient PARITY is checked with every ``*_dropout = 0`` and dropout itself by its
statistics and by directional derivatives.  Attention-probability dropout needs the tensor-core
attention path.  Dense-conv (non-depthwise) stacks train too (forward / input gradients on the tensor
cores as k shifted GEMMs, per-tap weight gradients on the exact-fp32 CUDA-core kernel).

Per optimizer step the operand forms of the weights (bf16 planes, transposed planes, the folded FFN-2 matrix, tap-major
depthwise weights) are derived once (``WeightCache``; after the first step by ONE batched launch) and shared by forward,
backward and -- with ``model.train_length_buckets = n`` -- by the n length-sorted sub-batches the step is cut into
(``forward_train_bucketed``: PAD rows beyond each bucket's longest utterance + conv halo are never computed; losses and
gradients equal the one-tensor step).
"""
import numpy as np
import torch

from .. import ops


def grad_of(p):
    """The gradient buffer of parameter p (allocated zeroed on first use; kernels accumulate into it)."""
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return p.grad


class WeightCache:
    """Tensors that depend on the parameters only -- bf16 planes, transposed copies, the folded FFN-2 matrix -- and the
    per-step accumulators of gradients that take a second kernel to reach the parameters (folded FFN-2 weights, transposed
    depthwise taps).  Built on first use after the parameters changed, shared by forward and backward and by every length
    bucket of the step; `flush()` at the end of a backward pass runs the deferred gradient kernels once."""

    batched_prep = True   # refill the planes of every GEMM weight with ONE launch per step (see _prefill)

    def __init__(self):
        self.key, self.store, self.pending = None, {}, {}
        self.seen, self.plan = {}, None

    def validate(self, model):
        key = (ops.WEIGHTS_EPOCH, model.compute_mode, tuple(p._version for p in model.parameters()))
        if key != self.key:
            seen, self.seen = self.seen, {}
            self.key, self.store, self.pending = key, {}, {}
            if self.batched_prep and seen and model.compute_mode != "simt":
                self._prefill(seen)
        # no backward pass is in flight when a forward starts (every backward ends in flush()): accumulators left behind
        # by a backward that raised must not leak into the next one
        self.pending = {}

    def get(self, tag, w, make):
        """w: a parameter or a view of one (its storage outlives the cache entry, so the address identifies it)"""
        k = (tag, w.data_ptr(), tuple(w.shape))
        v = self.store.get(k)
        if v is None:
            v = self.store[k] = make()
        if tag in ("planes", "planesT"):
            self.seen[k] = w
        return v

    def _prefill(self, seen):
        """The weights changed (optimizer step): every (planes, planesT) entry the LAST step asked for is rebuilt by one
        lfs2_weight_planes_batched launch into persistent buffers, instead of one split / transpose + split launch per
        matrix as the step reaches it.  Only sources that are parameters (or views of parameters) take part: tensors the
        cache derives itself (the folded FFN-2 matrix) do not exist yet for the new weights and stay on the lazy path."""
        want = {}
        for (tag, ptr, shape), w in seen.items():
            derived = not isinstance(w, torch.nn.Parameter) and w._base is None
            if derived or w.dim() != 2 or w.dtype != torch.float32 or not w.is_contiguous() or not w.is_cuda:
                continue
            e = want.setdefault((ptr, shape), [w, False, False])
            e[1 if tag == "planes" else 2] = True
        if not want:
            return
        sources = [(w, p, t) for w, p, t in want.values()]
        sig = tuple((w.data_ptr(), tuple(w.shape), p, t) for w, p, t in sources)
        if self.plan is None or self.plan.signature != sig:
            self.plan = ops.WeightPrepPlan(sources)
        self.plan.run()
        for (w, p, t), pl, pt in zip(sources, self.plan.planes, self.plan.planes_t):
            k = (w.data_ptr(), tuple(w.shape))
            if pl is not None:
                self.store[("planes",) + k] = pl
                self.seen[("planes",) + k] = w
            if pt is not None:
                self.store[("planesT",) + k] = pt
                self.seen[("planesT",) + k] = w

    def accumulator(self, tag, w, make, finish):
        """-> the step's accumulator for (tag, w), created by make(); finish(acc) runs once in flush()"""
        k = (tag, w.data_ptr(), tuple(w.shape))
        acc = self.pending.get(k)
        if acc is None:
            acc = make()
            self.pending[k] = (acc, finish)
            return acc
        return acc[0]

    def flush(self):
        pending, self.pending = self.pending, {}
        for acc, finish in pending.values():
            finish(acc)


class Engine:
    """GEMM dispatch of one train step: exact-fp32 CUDA-core kernels ("simt") or the tcgen05 kernels
    on bf16 hi/lo split operands ("fp32": 3 passes, "bf16": 1 pass; fp32 accumulation either way)."""

    def __init__(self, mode, cache=None):
        self.mode = mode
        self.cache = cache if cache is not None else WeightCache()
        self.tc = mode != "simt"
        self.npass = 3 if mode == "fp32" else 1
        # dropout: one 64-bit seed per forward pass from torch's CPU generator (torch.manual_seed makes runs
        # reproducible), one site number per dropout call; the backward replays (seed, site)
        self.seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        self.site = 0

    def site_token(self, p):
        """replay token of a dropout site that a kernel applies itself (fused), None if p == 0"""
        if p <= 0:
            return None
        self.site += 1
        return (p, self.seed, self.site)

    def dropout_(self, x, p):
        """in-place dropout; returns the replay token for the backward pass (None if p == 0)"""
        if p <= 0:
            return None
        self.site += 1
        tok = (p, self.seed, self.site)
        ops.dropout_(x, *tok)
        return tok

    def dropout_any(self, x, p):
        """dropout of an fp32 tensor (in place) or of Planes (new planes) -> (result, replay token)"""
        if not isinstance(x, ops.Planes):
            return x, self.dropout_(x, p)
        if p <= 0:
            return x, None
        self.site += 1
        tok = (p, self.seed, self.site)
        return ops.dropout_planes(x, *tok), tok

    @staticmethod
    def dropout_bwd(dy, tok, inplace=True):
        """gradient through a dropout site: same mask, same 1/(1-p) scale"""
        if tok is None:
            return dy
        return ops.dropout_(dy, *tok, out=None if inplace else torch.empty_like(dy))

    def _planes(self, x):
        """hi/lo planes of an fp32 tensor, split once and remembered on the tensor (activations are used by
        the forward GEMM and again by the weight-gradient GEMM)"""
        if isinstance(x, ops.Planes):
            return x
        p = getattr(x, "_lfs2_planes", None)
        if p is None:
            p = ops.split_bf16(x.contiguous())
            if not isinstance(x, torch.nn.Parameter):
                ops.attach_planes(x, p)
        return p

    def wplanes(self, w):
        """planes of a weight matrix (a parameter, a view of one, or a tensor the cache itself holds), once per step"""
        return self.cache.get("planes", w, lambda: ops.split_bf16(w.contiguous()))

    def wt(self, w):
        """(rows, cols) weight -> its fp32 transpose, once per step"""
        return self.cache.get("T", w, lambda: ops.transpose(w))

    def wplanes_t(self, w):
        return self.cache.get("planesT", w, lambda: ops.split_bf16(self.wt(w)))

    def folded(self, pw2, gc):
        """(w_eff, b_eff) of conv2.1 . conv2.0 (grouped 1x1), once per step"""
        return self.cache.get("fold", pw2.weight, lambda: ops.fold_pw(_mat(pw2.weight), _mat(gc.weight), gc.bias, pw2.bias))

    def dw_taps(self, conv, flipped=False):
        """(k, d) tap-major copy of a depthwise Conv1d weight (d, 1, k); flipped: reversed taps (input gradient)"""
        if flipped:
            return self.cache.get("dwTf", conv.weight, lambda: self.dw_taps(conv).flip(0).contiguous())
        return self.cache.get("dwT", conv.weight, lambda: ops.transpose(_dwmat(conv.weight)))

    def attention_fwd(self, qkv, kpm, nhead, p_drop=0.0):
        """-> (ctx, saved)"""
        d = qkv.shape[-1] // 3
        if isinstance(qkv, ops.Planes) or (self.tc and (d // nhead) % 32 == 0):
            drop = None
            if p_drop > 0:
                self.site += 1
                drop = (p_drop, self.seed, self.site)
            ctx, p, lse = ops.attention_mat_fwd(self._planes(qkv), kpm, nhead, npass=self.npass, drop=drop)
            return ctx, {"p": p, "drop": drop}
        if p_drop > 0:
            raise NotImplementedError(
                "attention-probability dropout needs the tensor-core attention (compute_mode 'fp32' or 'bf16' and "
                "head_dim % 32 == 0); the CUDA-core kernel implements p = 0 only")
        ctx, lse = ops.attention_lse(qkv, kpm, nhead)
        return ctx, {"lse": lse}

    def attention_bwd(self, qkv, ctx, dctx, saved, kpm, nhead):
        if "p" in saved:
            return ops.attention_mat_bwd(self._planes(qkv), saved["p"], ctx, dctx, nhead, npass=self.npass,
                                         drop=saved["drop"])
        return ops.attention_bwd(qkv, ctx, dctx, saved["lse"], kpm, nhead)

    def linear(self, x, w, b, relu=False, tag=None, planes_out=False):
        """x (..., k) . w (n, k)^T + b.  planes_out: on the tensor-core path return bf16 hi/lo Planes instead of an
        fp32 tensor (for results that only feed further tensor-core kernels: no fp32 copy, no split pass)."""
        n, k = w.shape
        if self.tc and k % 32 == 0 and n % 16 == 0:
            return ops.gemm_tc(self._planes(x), self.wplanes(w), b, relu=relu, npass=self.npass, tag=tag,
                               out="planes" if planes_out else "f32")
        return ops.linear(x if not isinstance(x, ops.Planes) else ops.merge_planes(x), w, b, relu=relu, tag=tag)

    def dwconv(self, x, wt, bias):
        """depthwise conv whose result only feeds GEMMs: Planes on the tensor-core path, fp32 otherwise"""
        if self.tc and x.shape[-1] % 32 == 0:
            return ops.dwconv1d_planes(x, wt, bias, out="planes")
        return ops.dwconv1d(x, wt, bias)

    def dgrad(self, dy, w, tag=None):
        """dy (..., n) . w (n, k) -> (..., k): the layer-input gradient of y = x . w^T"""
        n, k = w.shape
        if self.tc and n % 32 == 0 and k % 16 == 0:
            return ops.gemm_tc(self._planes(dy), self.wplanes_t(w), None, npass=self.npass, tag=tag)
        return ops.linear(dy if not isinstance(dy, ops.Planes) else ops.merge_planes(dy), self.wt(w), None, tag=tag)

    # -- dense Conv1d(c -> n, k, "same") over (B, T, c), weight in the reference layout (n, c, k) -------------
    def conv(self, x, w, b, relu=False, tag=None):
        n, c, k = w.shape
        wp = w.permute(0, 2, 1).reshape(n, k * c).contiguous()           # tap-major repack (weights only)
        if self.tc and c % 32 == 0 and n % 16 == 0:
            return ops.gemm_tc(self._planes(x), ops.split_bf16(wp), b, taps=k, relu=relu, npass=self.npass, tag=tag)
        return ops.conv1d_dense(x, wp, b, k, relu=relu, tag=tag)

    def conv_dgrad(self, dy, w, tag=None):
        """input gradient = the same convolution with reversed taps and transposed channels"""
        n, c, k = w.shape
        wt = w.flip(2).permute(1, 2, 0).reshape(c, k * n).contiguous()
        if self.tc and n % 32 == 0 and c % 16 == 0:
            return ops.gemm_tc(self._planes(dy), ops.split_bf16(wt), None, taps=k, npass=self.npass, tag=tag)
        return ops.conv1d_dense(dy, wt, None, k, tag=tag)

    def conv_wgrad_(self, dw, db, dy, x, tag=None):
        """dw (n, c, k) += per-tap dy^T . shift(x); db += column sums of dy"""
        n, c, k = dw.shape
        if k == 1:
            self.wgrad_(dw.view(n, c), db, dy, x, tag=tag)
            return
        t = x.shape[1]
        dwp = torch.zeros(n, k * c, device=dy.device, dtype=torch.float32)
        for j in range(k):  # exact-fp32 CUDA-core kernel: rows of x shifted by the tap offset inside each utterance
            ops.gemm_tn_(dwp, dy, x, t=t, shift=j - (k - 1) // 2, col_offset=j * c)
        ops.add_(dw, dwp.view(n, k, c).permute(0, 2, 1).contiguous())
        if db is not None:
            ops.colsum_(db, dy)

    def wgrad_(self, dw, db, dy, x, tag=None):
        """dw (n, k) += dy^T . x ; db (n) += column sums of dy"""
        n, k = dw.shape
        if self.tc and ops.wgrad_tc_ok(n, k):
            ops.gemm_wgrad_tc_(dw, self._planes(dy), self._planes(x), npass=self.npass, tag=tag)
        else:
            dy = ops.merge_planes(dy) if isinstance(dy, ops.Planes) else dy
            ops.gemm_tn_(dw, dy, ops.merge_planes(x) if isinstance(x, ops.Planes) else x)
        if db is not None:
            ops.colsum_(db, ops.merge_planes(dy) if isinstance(dy, ops.Planes) else dy)


def _mat(w):
    """(n, k, 1) pointwise-conv weight or its gradient as an (n, k) matrix view"""
    return w.view(w.shape[0], w.shape[1])


def _dwmat(w):
    """(d, 1, k) depthwise-conv weight or its gradient as a (d, k) matrix view"""
    return w.view(w.shape[0], w.shape[2])


# ------------------------------------------------------------------------------------------
# FFTBlock (reference model.py:108-122)
def fft_fwd(L, E, x, kpm):
    p = L.p_drop
    sa = L.self_attn
    if L.depthwise:
        dwc, pw, gc, pw2 = L.conv1[0], L.conv1[1], L.conv2[0], L.conv2[1]
        if gc.kernel_size[0] != 1:
            raise NotImplementedError("grouped conv2.0 with kernel > 1")
    s = {"x": x, "kpm": kpm}
    d_model = x.shape[-1]
    qkv_planes = E.tc and (d_model // L.nhead) % 32 == 0   # consumed only by the attention GEMMs and the wgrad
    s["qkv"] = E.linear(x, sa.in_proj_weight, sa.in_proj_bias, tag="qkv_gemm", planes_out=qkv_planes)
    s["ctx"], s["att"] = E.attention_fwd(s["qkv"], kpm, L.nhead, p)
    a = E.linear(s["ctx"], sa.out_proj.weight, sa.out_proj.bias, tag="out_proj_gemm")
    s["drop1"] = E.site_token(p)                                          # dropout1 (model.py:114), fused into the LN kernel
    x1, s["z1"], s["st1"] = ops.add_layernorm_train(x, a, L.norm1.weight, L.norm1.bias, L.eps, drop=s["drop1"])
    s["x1"] = x1
    if L.depthwise:
        s["u"] = E.dwconv(x1, E.dw_taps(dwc), dwc.bias)                   # taps as (k, d)
        # the F-wide activation only feeds GEMMs (FFN-2 forward, FFN-1 weight gradient) and the ReLU mask of the
        # backward: on the tensor-core path it exists as bf16 planes only (no fp32 copy, no split pass)
        s["v"] = E.linear(s["u"], _mat(pw.weight), pw.bias, relu=True, tag="ffn1_gemm", planes_out=E.tc)
        s["v"], s["dropv"] = E.dropout_any(s["v"], p)                     # dropout after ReLU (model.py:120)
        w_eff, b_eff = E.folded(pw2, gc)
        y = E.linear(s["v"], w_eff, b_eff, tag="ffn2_gemm")
    else:                                                                  # dense convolutions (model.py:95-106)
        s["v"] = E.conv(x1, L.conv1.weight, L.conv1.bias, relu=True, tag="ffn1_gemm")
        s["dropv"] = E.dropout_(s["v"], p)
        y = E.conv(s["v"], L.conv2.weight, L.conv2.bias, tag="ffn2_gemm")
    s["drop2"] = E.site_token(p)                                          # dropout2 (model.py:115), fused into the LN kernel
    x2, s["z2"], s["st2"] = ops.add_layernorm_train(x1, y, L.norm2.weight, L.norm2.bias, L.eps, drop=s["drop2"])
    return x2, s


def fft_bwd(L, E, s, dx2, dx2b=None):
    """gradient of the block's output, given as dx2 (+ dx2b: the two summands of the residual join of the block above)
    -> (dx, dxb): the gradient of the block's input as two summands as well -- the joins are added on load by the next
    LayerNorm backward (ops.layernorm_bwd dy2) instead of by add kernels; callers at a stack's end add them up"""
    dev = dx2.device
    dz2 = ops.layernorm_bwd(dx2, s["z2"], s["st2"], L.norm2.weight, grad_of(L.norm2.weight), grad_of(L.norm2.bias),
                            drop=s["drop2"], dy2=dx2b)
    dz2, dy = dz2 if s["drop2"] else (dz2, dz2)                          # dy = dropout2's backward of dz2 (same kernel)
    if L.depthwise:
        dx1 = _ffn_bwd_depthwise(L, E, s, dy, dev)
    else:
        dv = E.conv_dgrad(dy, L.conv2.weight, tag="ffn2_dgrad")
        E.conv_wgrad_(grad_of(L.conv2.weight), grad_of(L.conv2.bias), dy, s["v"], tag="ffn2_wgrad")
        ops.relu_bwd_(dv, s["v"], scale=1.0 / (1.0 - s["dropv"][0]) if s["dropv"] else 1.0)
        dx1 = E.conv_dgrad(dv, L.conv1.weight, tag="ffn1_dgrad")
        E.conv_wgrad_(grad_of(L.conv1.weight), grad_of(L.conv1.bias), dv, s["x1"], tag="ffn1_wgrad")
    return _attn_bwd(L, E, s, dx1, dz2)                                  # (dz2: the residual around the FFN)


def _ffn_bwd_depthwise(L, E, s, dy, dev):
    """conv1 = depthwise(k1) + pointwise, conv2 = grouped 1x1 + pointwise (folded): returns d(x1) without the residual"""
    dwc, pw, gc, pw2 = L.conv1[0], L.conv1[1], L.conv2[0], L.conv2[1]
    w_eff, _ = E.folded(pw2, gc)
    dv = E.dgrad(dy, w_eff, tag="ffn2_dgrad")

    def finish(acc):  # chain rule from the folded matrix back to the four reference tensors, once per step
        ops.fold_pw_bwd_(acc[0], acc[1], _mat(pw2.weight), _mat(gc.weight), gc.bias, _mat(grad_of(pw2.weight)),
                         _mat(grad_of(gc.weight)), grad_of(gc.bias), grad_of(pw2.bias))

    dw_eff, db_eff = E.cache.accumulator(
        "d_fold", pw2.weight,
        lambda: (torch.zeros_like(w_eff), torch.zeros(w_eff.shape[0], device=dev, dtype=torch.float32)), finish)
    E.wgrad_(dw_eff, db_eff, dy, s["v"], tag="ffn2_wgrad")
    # ReLU + the dropout behind it in one pass: s["v"] is the dropped activation, so v > 0 is both masks
    scale = 1.0 / (1.0 - s["dropv"][0]) if s["dropv"] else 1.0
    if E.tc:   # the masked gradient as planes (its two consumers are GEMMs) + the bias gradient, one kernel
        dv = ops.relu_bwd_planes(dv, s["v"], scale=scale, db=grad_of(pw.bias))
        du = E.dgrad(dv, _mat(pw.weight), tag="ffn1_dgrad")
        E.wgrad_(_mat(grad_of(pw.weight)), None, dv, s["u"], tag="ffn1_wgrad")
    else:
        ops.relu_bwd_(dv, s["v"], scale=scale)
        du = E.dgrad(dv, _mat(pw.weight), tag="ffn1_dgrad")
        E.wgrad_(_mat(grad_of(pw.weight)), grad_of(pw.bias), dv, s["u"], tag="ffn1_wgrad")
    return _dwconv_bwd(E, dwc, du, s["x1"])


def _attn_bwd(L, E, s, dx1, dx1b):
    sa = L.self_attn
    dz1 = ops.layernorm_bwd(dx1, s["z1"], s["st1"], L.norm1.weight, grad_of(L.norm1.weight), grad_of(L.norm1.bias),
                            drop=s["drop1"], dy2=dx1b)
    dz1, da = dz1 if s["drop1"] else (dz1, dz1)
    dctx = E.dgrad(da, sa.out_proj.weight, tag="out_proj_dgrad")
    E.wgrad_(grad_of(sa.out_proj.weight), grad_of(sa.out_proj.bias), da, s["ctx"], tag="out_proj_wgrad")
    dqkv = E.attention_bwd(s["qkv"], s["ctx"], dctx, s["att"], s["kpm"], L.nhead)
    dx = E.dgrad(dqkv, sa.in_proj_weight, tag="qkv_dgrad")
    E.wgrad_(grad_of(sa.in_proj_weight), grad_of(sa.in_proj_bias), dqkv, s["x"], tag="qkv_wgrad")
    return dx, dz1                                                       # (dz1: the residual around the attention)


def _dwconv_bwd(E, conv, du, x_in):
    """depthwise Conv1d backward: returns d(input); accumulates weight (d,1,k) and bias gradients (the tap-major
    weight gradient of the step is transposed into the parameter's layout once, in the cache's flush)"""
    taps = E.dw_taps(conv)
    k, d = taps.shape
    zero_b = E.cache.get("zero_bias", conv.bias, lambda: torch.zeros(d, device=du.device, dtype=torch.float32))
    dx = ops.dwconv1d(du, E.dw_taps(conv, flipped=True), zero_b)         # correlation with the reversed taps
    dwt = E.cache.accumulator("d_dwT", conv.weight, lambda: torch.zeros_like(taps),
                              lambda acc: ops.add_(_dwmat(grad_of(conv.weight)), ops.transpose(acc)))
    ops.dwconv1d_bwd_w_(dwt, grad_of(conv.bias), du, x_in)
    return dx


# ------------------------------------------------------------------------------------------
# VariancePredictor (reference model.py:482-561)
def vp_fwd(P, E, x, mask):
    layers = []
    z = x
    for layer in P.layers:
        conv, ln = layer.layers[0].module, layer.layers[2]
        if layer.depthwise:
            u = E.dwconv(z, E.dw_taps(conv[0]), conv[0].bias)
            h = E.linear(u, _mat(conv[1].weight), conv[1].bias, relu=True, tag="predictor_pw_gemm")
        else:
            u = None
            h = E.conv(z, conv.weight, conv.bias, relu=True, tag="predictor_conv_gemm")
        zo, _, st = ops.add_layernorm_train(h, None, ln.weight, ln.bias, ln.eps)
        layers.append({"x": z, "u": u, "h": h, "st": st, "drop": E.dropout_(zo, layer.layers[3].p)})
        z = zo
    out = ops.rowdot_mask(z, P.linear.weight, P.linear.bias, mask)
    return out, {"layers": layers, "z": z, "mask": mask}


def vp_bwd(P, E, s, dout):
    dz = ops.rowdot_mask_bwd(dout.contiguous(), s["z"], P.linear.weight, s["mask"], grad_of(P.linear.weight),
                             grad_of(P.linear.bias))
    for layer, sl in zip(reversed(list(P.layers)), reversed(s["layers"])):
        conv, ln = layer.layers[0].module, layer.layers[2]
        E.dropout_bwd(dz, sl["drop"])
        dh = ops.layernorm_bwd(dz, sl["h"], sl["st"], ln.weight, grad_of(ln.weight), grad_of(ln.bias))
        if layer.depthwise and E.tc:
            dh = ops.relu_bwd_planes(dh, sl["h"], db=grad_of(conv[1].bias))
            du = E.dgrad(dh, _mat(conv[1].weight), tag="predictor_dgrad")
            E.wgrad_(_mat(grad_of(conv[1].weight)), None, dh, sl["u"], tag="predictor_wgrad")
            dz = _dwconv_bwd(E, conv[0], du, sl["x"])
            continue
        ops.relu_bwd_(dh, sl["h"])
        if layer.depthwise:
            du = E.dgrad(dh, _mat(conv[1].weight), tag="predictor_dgrad")
            E.wgrad_(_mat(grad_of(conv[1].weight)), grad_of(conv[1].bias), dh, sl["u"], tag="predictor_wgrad")
            dz = _dwconv_bwd(E, conv[0], du, sl["x"])
        else:
            dz = E.conv_dgrad(dh, conv.weight, tag="predictor_dgrad")
            E.conv_wgrad_(grad_of(conv.weight), grad_of(conv.bias), dh, sl["x"], tag="predictor_wgrad")
    return dz


# ------------------------------------------------------------------------------------------
# whole model (reference fastspeech2.py:636-784, teacher-forced branch)
def forward_train(M, targets, frames=None):
    """-> (result dict with the reference's keys, saved state for backward_train).
    frames = (l, cap): length of the LengthRegulator output (l >= cap, frames in [cap, l) are PAD) for a length bucket
    whose tensor ends before / beyond its own longest utterance (forward_train_bucketed); None = the batch's own."""
    hp = M.hparams
    dev = M.device
    cache = M.__dict__.setdefault("_train_weight_cache", WeightCache())
    cache.validate(M)
    E = Engine(M.compute_mode, cache)
    va = M.variance_adaptor
    phones = targets["phones"].to(dev, non_blocking=True).contiguous()
    dvec = targets["speaker"].to(dev, dtype=torch.float32, non_blocking=True).contiguous()
    pe = M.positional_encoding.pe
    S = {"E": E, "phones": phones, "dvec": dvec}

    spk = ops.speaker_proj(dvec, M.speaker_embedding.projection.weight, M.speaker_embedding.projection.bias)
    S["spk"] = spk
    p_pe = M.positional_encoding.dropout.p
    if p_pe > 0:  # dropout sits between "+ PE" and "+ spk" (fastspeech2.py:653-660): split the fused front end
        zero_spk = torch.zeros_like(spk)
        zero_pe = torch.zeros(1, max(phones.shape[1], 1), pe.shape[-1], device=dev, dtype=torch.float32)
        x, src_mask = ops.embed_pe_spk(phones, M.phone_embedding.weight, pe, zero_spk)
        S["drop_enc"] = E.dropout_(x, p_pe)
        ops.add_pe_spk_(x, zero_pe, spk)
    else:
        S["drop_enc"] = None
        x, src_mask = ops.embed_pe_spk(phones, M.phone_embedding.weight, pe, spk)
    S["enc"] = []
    for L in M.encoder.layers:
        x, s = fft_fwd(L, E, x, src_mask)
        S["enc"].append(s)

    S["priors"] = []
    if len(hp.priors):  # per-utterance prior embeddings added to the encoder output (fastspeech2.py:687-692)
        zero_pe = torch.zeros(1, max(x.shape[1], 1), x.shape[2], device=dev)
        for prior in hp.priors:
            term, idx = M.prior_embeddings[prior].term(targets[f"priors_{prior}"])
            ops.add_pe_spk_(x, zero_pe, term)
            S["priors"].append((prior, term, idx))
    dur_pred, S["dp"] = vp_fwd(va.duration_predictor, E, x, src_mask)
    np.random.uniform(0, 1)  # the reference draws the teacher-forcing coin here (model.py:272); tf_ratio = 1
    result = {}
    S["phone_vars"] = []
    for i, var in enumerate(va.variances):  # phone-level variances act on the encoder output (model.py:277-294)
        if va.variance_levels[i] != "phone":
            continue
        enc = va.encoders[var]
        pred, s_vp = vp_fwd(enc.predictor, E, x, src_mask)
        tgt = targets[f"variances_{var}"].to(dev, dtype=torch.float32)[:, : x.shape[1]].contiguous()
        x, idx = ops.bucket_embed_add_oop(x, tgt, enc.std, enc.mean, enc.bins, enc.embedding.weight)
        S["phone_vars"].append((var, s_vp, idx))
        result[f"variances_{var}"] = pred
    duration = targets["duration"].to(dev)
    x, tgt_mask, S["cum"] = ops.length_regulate_train(x, duration, va.max_length, frames=frames)
    S["vars"] = []
    for i, var in enumerate(va.variances):
        if va.variance_levels[i] != "frame":
            continue
        enc = va.encoders[var]
        pred, s_vp = vp_fwd(enc.predictor, E, x, tgt_mask)
        tgt = targets[f"variances_{var}"].to(dev, dtype=torch.float32)[:, : x.shape[1]].contiguous()
        x, idx = ops.bucket_embed_add_oop(x, tgt, enc.std, enc.mean, enc.bins, enc.embedding.weight)
        S["vars"].append((var, s_vp, idx))
        result[f"variances_{var}"] = pred

    if p_pe > 0:  # fastspeech2.py:703-707
        zero_pe = torch.zeros(1, max(x.shape[1], 1), pe.shape[-1], device=dev, dtype=torch.float32)
        ops.add_pe_spk_(x, pe, torch.zeros_like(spk))
        S["drop_dec"] = E.dropout_(x, p_pe)
        ops.add_pe_spk_(x, zero_pe, spk)
    else:
        S["drop_dec"] = None
        x = ops.add_pe_spk_(x, pe, spk)
    S["dec"] = []
    for L in M.decoder.layers:
        x, s = fft_fwd(L, E, x, tgt_mask)
        S["dec"].append(s)
    S["dec_out"] = x
    result["mel"] = E.linear(x, M.linear.weight, M.linear.bias, tag="mel_linear")
    result["duration_prediction"] = dur_pred
    result["duration_rounded"] = duration
    result["src_mask"] = src_mask
    result["tgt_mask"] = tgt_mask
    return result, S


def backward_train(M, S, dmel, ddur, dvars, flush=True):
    """dmel (B,Tm,80), ddur (B,Tp), dvars {var: (B,Tm)} (any may be None) -> accumulates every parameter gradient.
    flush=False leaves the deferred gradient kernels (WeightCache.flush) to the caller: one run per step, not per bucket"""
    E = S["E"]
    va = M.variance_adaptor
    dev = M.device
    M.rehome_gradients()  # gradients dropped by zero_grad(set_to_none=True) become views of the flat buffer again
    bsz = S["phones"].shape[0]
    d = M.hparams.encoder_hidden
    dspk = torch.zeros(bsz, d, device=dev, dtype=torch.float32)

    if dmel is not None:
        dmel = dmel.contiguous()
        dx = E.dgrad(dmel, M.linear.weight, tag="mel_dgrad")
        E.wgrad_(grad_of(M.linear.weight), grad_of(M.linear.bias), dmel, S["dec_out"], tag="mel_wgrad")
        dxb = None
        for L, s in zip(reversed(list(M.decoder.layers)), reversed(S["dec"])):
            dx, dxb = fft_bwd(L, E, s, dx, dxb)
        if dxb is not None:
            ops.add_(dx, dxb)
        ops.sum_over_time_(dspk, dx)                                     # "+ spk" at Tm (fastspeech2.py:707)
        E.dropout_bwd(dx, S["drop_dec"])
    else:
        dx = torch.zeros_like(S["dec_out"])
    for var, s_vp, idx in reversed(S["vars"]):
        enc = va.encoders[var]
        ops.embedding_bwd_(grad_of(enc.embedding.weight), dx, idx)
        g = dvars.get(var)
        if g is not None:
            ops.add_(dx, vp_bwd(enc.predictor, E, s_vp, g))
    dx = ops.length_regulate_bwd(dx, S["cum"])
    for var, s_vp, idx in reversed(S["phone_vars"]):
        enc = va.encoders[var]
        ops.embedding_bwd_(grad_of(enc.embedding.weight), dx, idx)
        g = dvars.get(var)
        if g is not None:
            ops.add_(dx, vp_bwd(enc.predictor, E, s_vp, g))
    if ddur is not None:
        ops.add_(dx, vp_bwd(va.duration_predictor, E, S["dp"], ddur))
    for prior, term, idx in S["priors"]:  # relu(emb[bucket]) broadcast over time: sum the gradient over t, mask, scatter
        g = torch.zeros(bsz, d, device=dev, dtype=torch.float32)
        ops.sum_over_time_(g, dx)
        ops.relu_bwd_(g, term)
        ops.embedding_bwd_(grad_of(M.prior_embeddings[prior].embedding.weight), g, idx)
    dxb = None
    for L, s in zip(reversed(list(M.encoder.layers)), reversed(S["enc"])):
        dx, dxb = fft_bwd(L, E, s, dx, dxb)
    if dxb is not None:
        ops.add_(dx, dxb)
    ops.sum_over_time_(dspk, dx)                                         # "+ spk" at Tp (fastspeech2.py:658)
    E.dropout_bwd(dx, S["drop_enc"])
    ops.embedding_bwd_(grad_of(M.phone_embedding.weight), dx, S["phones"], skip_idx=0)
    # speaker term: spk = relu(W . dvec + b)   (model.py:137-143)
    proj = M.speaker_embedding.projection
    ops.relu_bwd_(dspk, S["spk"])
    ops.gemm_tn_(grad_of(proj.weight), dspk, S["dvec"])
    ops.colsum_(grad_of(proj.bias), dspk)
    if flush:
        E.cache.flush()


# ------------------------------------------------------------------------------------------
# length-bucketed train step (SURVEY 8f N2 carried to the gradient path)
def plan_length_buckets(nphones, nframes, ngroups, tp, cap, h_enc, h_dec):
    """Host-side plan of the length-bucketed step.  nphones[i] / nframes[i]: phones (last valid + 1) and frames (sum of
    durations) of utterance i; tp: width of the collated phone tensor; cap: the LengthRegulator's max_length;
    h_enc / h_dec: conv halos of the encoder and decoder side.
    -> [(utterance indices, phones kept, (frames kept, LengthRegulator cap of the bucket))], longest bucket first.
    A bucket keeps its longest utterance + the halo, but never more than the full batch's own tensor would hold
    (that is where the reference's convolutions see their zero padding)."""
    bsz = len(nphones)
    cap = int(cap)
    l_full = min(max(nframes), cap)
    order = sorted(range(bsz), key=lambda i: (-nframes[i], -nphones[i]))
    ngroups = max(1, min(int(ngroups), bsz))
    per = (bsz + ngroups - 1) // ngroups
    plan = []
    for g0 in range(0, bsz, per):
        idx = order[g0:g0 + per]
        tp_g = min(tp, max(nphones[i] for i in idx) + h_enc)
        cap_g = min(max(nframes[i] for i in idx), cap)
        plan.append((idx, tp_g, (min(cap_g + h_dec, l_full), cap_g)))
    return plan


def forward_train_bucketed(M, targets, ngroups):
    """The teacher-forced forward of a ragged batch as `ngroups` length-sorted sub-batches, each padded only to ITS
    longest utterance plus the conv halo (`FastSpeech2._halos`), results scattered back into full-batch tensors.

    Exactness (same argument as `_forward_bucketed`): PAD rows are never attention keys and the loss masks them, so a
    PAD row reaches the loss only through the FFN / predictor convolutions of the layers above it; rows within the
    summed conv half-widths of an utterance's end are kept and hold exactly what the reference computes there
    (PE[t] + speaker term + embedding of the collated target's padding value, ...), rows farther out have a zero
    gradient and feed nothing.  Losses and parameter gradients therefore equal the un-bucketed step up to fp32
    summation order; result positions beyond a bucket's own tensor (all masked by the loss) come back as zeros instead of
    the reference's PAD-row values.
    -> (full-batch result dict, [(index tensor, tp_g, l_g, saved state), ...])"""
    dev, hp, va = M.device, M.hparams, M.variance_adaptor
    phones_all, dur_all = targets["phones"], targets["duration"]
    bsz, tp = phones_all.shape
    pos = torch.arange(1, tp + 1, device=phones_all.device)
    nphones = ((phones_all != 0) * pos).amax(1)                                   # last valid phone + 1
    nphones, nframes = torch.stack([nphones.to(torch.int64),
                                    dur_all.sum(1).to(nphones.device, torch.int64)]).tolist()    # ONE read-back
    cap = int(va.max_length)
    l_full = min(max(nframes), cap)
    h_enc = sum(layer.halo() for layer in M.encoder.layers) + max(
        [va.duration_predictor.halo()] + [va.encoders[v].predictor.halo() for i, v in enumerate(va.variances)
                                          if va.variance_levels[i] == "phone"])
    h_dec = max([sum(layer.halo() for layer in M.decoder.layers)] +
                [va.encoders[v].predictor.halo() for i, v in enumerate(va.variances) if va.variance_levels[i] == "frame"])
    parts = []
    for idx, tp_g, frames in plan_length_buckets(nphones, nframes, ngroups, tp, cap, h_enc, h_dec):
        it = {}

        def take(v, it=it, idx=idx):
            if not (torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == bsz):
                return v
            if v.device not in it:
                it[v.device] = torch.tensor(idx, device=v.device)
            return v[it[v.device]]

        sub = {k: take(v) for k, v in targets.items()}
        sub["phones"] = sub["phones"][:, :tp_g].contiguous()
        sub["duration"] = sub["duration"][:, :tp_g].contiguous()
        r, s = forward_train(M, sub, frames=frames)
        parts.append((torch.tensor(idx, device=dev), tp_g, frames[0], r, s))
    out = {"mel": torch.zeros(bsz, l_full, hp.n_mels, device=dev),
           "duration_prediction": torch.zeros(bsz, tp, device=dev),
           "duration_rounded": dur_all.to(dev),
           "src_mask": phones_all.to(dev) == 0,
           "tgt_mask": torch.ones(bsz, l_full, device=dev, dtype=torch.bool)}
    for i, v in enumerate(va.variances):
        out[f"variances_{v}"] = torch.zeros(bsz, tp if va.variance_levels[i] == "phone" else l_full, device=dev)
    saved = []
    for it, tp_g, l_g, r, s in parts:
        out["mel"][it, :l_g] = r["mel"]
        out["tgt_mask"][it, :l_g] = r["tgt_mask"]
        out["duration_prediction"][it, :tp_g] = r["duration_prediction"]
        for i, v in enumerate(va.variances):
            w = tp_g if va.variance_levels[i] == "phone" else l_g
            out[f"variances_{v}"][it, :w] = r[f"variances_{v}"]
        saved.append((it, tp_g, l_g, s))
    return out, saved


def backward_train_bucketed(M, saved, dmel, ddur, dvars):
    """gradients of the full-batch results -> one backward_train per length bucket (parameter gradients accumulate)"""
    va = M.variance_adaptor
    levels = dict(zip(va.variances, va.variance_levels))
    for it, tp_g, l_g, s in saved:
        dv = {v: g[it, :(tp_g if levels[v] == "phone" else l_g)].contiguous() for v, g in dvars.items() if g is not None}
        backward_train(M, s, None if dmel is None else dmel[it, :l_g].contiguous(),
                       None if ddur is None else ddur[it, :tp_g].contiguous(), dv, flush=False)
    if saved:
        saved[0][3]["E"].cache.flush()


class ForwardTrainFn(torch.autograd.Function):
    """autograd boundary of the train step: tensors out, gradients in; everything between is kernels
    of liblfs2.so.  ``anchor`` is a dummy scalar that requires grad so that autograd calls backward;
    parameter gradients are accumulated into p.grad directly (torch's own semantics for leaves)."""

    @staticmethod
    def forward(ctx, anchor, model, targets, keys):
        nb = int(getattr(model, "train_length_buckets", 1))
        ctx.bucketed = nb > 1 and targets["phones"].shape[0] > 1
        result, saved = forward_train_bucketed(model, targets, nb) if ctx.bucketed else forward_train(model, targets)
        ctx.model, ctx.saved, ctx.keys = model, saved, keys
        model._last_train_result = result
        return tuple(result[k] for k in keys)

    @staticmethod
    def backward(ctx, *grads):
        g = dict(zip(ctx.keys, grads))
        dvars = {k[len("variances_"):]: v for k, v in g.items() if k.startswith("variances_")}
        (backward_train_bucketed if ctx.bucketed else backward_train)(ctx.model, ctx.saved, g.get("mel"),
                                                                      g.get("duration_prediction"), dvars)
        ctx.saved = None
        return None, None, None, None
