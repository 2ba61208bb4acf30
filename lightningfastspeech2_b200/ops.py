"""Tensor-level wrappers over the C ABI (include/lfs2.h).

PyTorch is plumbing here: it owns device memory and the stream; every function below
validates shapes/dtypes, allocates outputs with torch.empty and hands raw pointers to the
hand-written kernels.  Nothing in this file computes on the host or falls back to ATen.
"""
import ctypes

import torch

from . import _lib

LN_EPS = 1e-5

# bumped whenever a kernel rewrites parameters through raw pointers (fused optimizer, re-homed flat
# buffers): tensor._version cannot see those writes, so weight-pack caches key on this as well
WEIGHTS_EPOCH = 0

# When bench.py sets PROFILE = {} every launch is bracketed by CUDA events on the launching
# stream and tagged with its algorithmic flops / HBM bytes (SURVEY 8d definitions).
PROFILE = None


def _launch(cname, *args, tag=None, flops=0.0, nbytes=0.0, passes=1):
    """passes: tensor-core MMA passes issued per algorithmic product (3 for split-bf16 operands, 1 otherwise)"""
    fn = getattr(_lib.lib(), cname)
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        PROFILE.setdefault(tag or cname, []).append((e0, e1, float(flops), float(nbytes), float(flops) * passes))
    else:
        rc = fn(*args)
    _lib.CALLS += 1
    _lib.check(rc, cname)


def measured_peaks():
    """(HBM GB/s, dense bf16 TFLOP/s sustained) from MEASURED_PEAKS.json at the repo root (driver-written), else the
    fallback figures of the profiling recipe"""
    import json
    import os

    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"]))
    except Exception:  # noqa: BLE001
        return 6650.0, 1400.0


def collect_profile(hbm_gbs=None, tensor_tflops=None):
    """{tag: {ms, launches, flops, bytes, bound}} summed over the recorded launches; `bound` is
    whichever of (bytes / HBM peak, flops / tensor peak) is the longer time (peaks: MEASURED_PEAKS.json)."""
    if hbm_gbs is None or tensor_tflops is None:
        hbm_gbs, tensor_tflops = measured_peaks()
    torch.cuda.synchronize()
    out = {}
    for tag, recs in (PROFILE or {}).items():
        ms = sum(r[0].elapsed_time(r[1]) for r in recs)
        fl = sum(r[2] for r in recs)
        by = sum(r[3] for r in recs)
        issued = sum(r[4] for r in recs)   # flops the tensor pipe executes (x3 for split-bf16 products)
        bound = "tensor" if issued / (tensor_tflops * 1e12) > by / (hbm_gbs * 1e9) else "hbm"
        out[tag] = {"ms": ms, "launches": len(recs), "flops": fl, "bytes": by, "issued_flops": issued, "bound": bound}
    return out


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _s():
    # raw handle of torch's current stream on the current device (the public torch.cuda.current_stream() builds
    # a Stream object per call: ~5 us of host time on every launch)
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


def _chk(t, dtype, name, ndim=None):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise _lib.Lfs2Error(f"{name}: expected a CUDA tensor (no CPU fallback exists)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name}: expected {ndim} dims, got {tuple(t.shape)}")
    return t


def speaker_proj(dvec, w, b):
    _chk(dvec, torch.float32, "speaker", 2); _chk(w, torch.float32, "projection.weight", 2)
    out = torch.empty(dvec.shape[0], w.shape[0], device=dvec.device, dtype=torch.float32)
    _launch("lfs2_speaker_proj", _p(dvec), _p(w), _p(b), _p(out), dvec.shape[0], dvec.shape[1],
                                            w.shape[0], _s())
    return out


def embed_pe_spk(phones, emb, pe, spk):
    _chk(phones, torch.int64, "phones", 2); _chk(emb, torch.float32, "phone_embedding.weight", 2)
    b, t = phones.shape
    d = emb.shape[1]
    if t > pe.shape[-2]:
        raise ValueError(f"sequence length {t} exceeds the positional table ({pe.shape[-2]})")
    x = torch.empty(b, t, d, device=phones.device, dtype=torch.float32)
    mask = torch.empty(b, t, device=phones.device, dtype=torch.bool)
    _launch("lfs2_embed_pe_spk", _p(phones), _p(emb), _p(pe), _p(spk), _p(x), _p(mask), b, t, d,
                                            emb.shape[0], _s())
    return x, mask


def add_pe_spk_(x, pe, spk):
    _chk(x, torch.float32, "x", 3)
    b, t, d = x.shape
    if t > pe.shape[-2]:
        raise ValueError(f"sequence length {t} exceeds the positional table ({pe.shape[-2]})")
    _launch("lfs2_add_pe_spk", _p(x), _p(pe), _p(spk), b, t, d, _s(), nbytes=2 * x.numel() * 4)
    drop_planes(x)
    return x


def linear(a, w, bias, relu=False, out=None, tag=None):
    """a (..., k) . w (n, k)^T + bias -> (..., n)"""
    _chk(a, torch.float32, "linear input"); _chk(w, torch.float32, "linear weight", 2)
    k = a.shape[-1]
    m = a.numel() // k
    n = w.shape[0]
    if w.shape[1] != k:
        raise ValueError(f"linear: weight {tuple(w.shape)} does not match input features {k}")
    if out is None:
        out = torch.empty(*a.shape[:-1], n, device=a.device, dtype=torch.float32)
    _launch("lfs2_linear", _p(a), _p(w), _p(bias), _p(out), m, n, k, int(relu), _s(), tag=tag or f"linear_n{n}_k{k}",
            flops=2.0 * m * n * k, nbytes=4.0 * (m * k + n * k + m * n))
    return out


def conv1d_dense(x, wp, bias, ksize, relu=False, tag=None):
    """x (B,T,d), wp (n, ksize*d) tap-major -> (B,T,n)"""
    _chk(x, torch.float32, "conv input", 3); _chk(wp, torch.float32, "conv weight", 2)
    b, t, d = x.shape
    n = wp.shape[0]
    if wp.shape[1] != ksize * d:
        raise ValueError("conv1d_dense: packed weight shape mismatch")
    out = torch.empty(b, t, n, device=x.device, dtype=torch.float32)
    _launch("lfs2_conv1d_dense", _p(x), _p(wp), _p(bias), _p(out), b, t, d, n, ksize, int(relu), _s(),
            tag=tag or f"conv_dense_k{ksize}_n{n}", flops=2.0 * b * t * n * ksize * d,
            nbytes=4.0 * (b * t * d + n * ksize * d + b * t * n))
    return out


def dwconv1d(x, wt, bias):
    """x (B,T,d), wt (ksize, d) -> (B,T,d)"""
    _chk(x, torch.float32, "dwconv input", 3); _chk(wt, torch.float32, "dwconv weight", 2)
    b, t, d = x.shape
    out = torch.empty_like(x)
    _launch("lfs2_dwconv1d", _p(x), _p(wt), _p(bias), _p(out), b, t, d, wt.shape[0], _s(),
            flops=2.0 * b * t * d * wt.shape[0], nbytes=8.0 * b * t * d)
    return out


def attention(qkv, kpm, nhead):
    """qkv (B,T,3d) packed [q|k|v], kpm (B,T) bool True=PAD -> ctx (B,T,d)"""
    _chk(qkv, torch.float32, "qkv", 3)
    b, t, d3 = qkv.shape
    d = d3 // 3
    if kpm is not None:
        _chk(kpm, torch.bool, "key_padding_mask", 2)
    ctx = torch.empty(b, t, d, device=qkv.device, dtype=torch.float32)
    fl = 0.0
    if PROFILE is not None:  # algorithmic flops: every query row x the utterance's VALID keys
        nkeys = (~kpm).sum(1).double() if kpm is not None else torch.full((b,), float(t))
        fl = float(4.0 * d * t * nkeys.sum())
    _launch("lfs2_attention", _p(qkv), _p(kpm), _p(ctx), b, t, d, nhead, _s(), flops=fl,
            nbytes=4.0 * (qkv.numel() + ctx.numel()))
    return ctx


def add_layernorm(x, y, gamma, beta, eps=LN_EPS):
    _chk(x, torch.float32, "layernorm input")
    if y is not None:
        _chk(y, torch.float32, "layernorm residual")
    d = x.shape[-1]
    m = x.numel() // d
    out = torch.empty_like(x)
    _launch("lfs2_add_layernorm", _p(x), _p(y), _p(gamma), _p(beta), _p(out), m, d, eps, _s(),
            nbytes=4.0 * m * d * (3 if y is not None else 2))
    return out


def add_layernorm_planes(x, y, gamma, beta, eps=LN_EPS, row_limit=None):
    """LayerNorm(x + y): x Planes (residual stream), y fp32 tensor or None -> Planes.
    row_limit = (lengths int32 (B), extra, ...) on (B,T,d) operands: rows of the 128-row groups starting at or after
    lengths[b] + extra are neither read nor written (FastSpeech2.skip_pad_rows)."""
    _chk(x.hi, torch.bfloat16, "layernorm residual planes"); _chk(x.lo, torch.bfloat16, "layernorm residual planes")
    if y is not None:
        _chk(y, torch.float32, "layernorm branch")
    d = x.shape[-1]
    m = x.hi.numel() // d
    out = _empty_planes(tuple(x.shape), x.hi.device)
    if row_limit is None:
        _launch("lfs2_add_layernorm_planes", _p(x.hi), _p(x.lo), _p(y), _p(gamma), _p(beta), _p(out.hi), _p(out.lo), m, d,
                float(eps), _s(), tag="lfs2_add_layernorm", nbytes=4.0 * m * d * (3 if y is not None else 2))
        return out
    if x.hi.dim() != 3:
        raise ValueError("add_layernorm_planes: row limits need (B,T,d) operands")
    b, t = x.hi.shape[:2]
    frac = _limited_fraction(row_limit, t)
    _launch("lfs2_add_layernorm_planes_limited", _p(x.hi), _p(x.lo), _p(y), _p(gamma), _p(beta), _p(out.hi), _p(out.lo), b, t,
            d, float(eps), _p(row_limit[0]), int(row_limit[1]), _s(), tag="lfs2_add_layernorm",
            nbytes=4.0 * m * d * (3 if y is not None else 2) * frac)
    return out


def rowdot_mask(z, w, bias, mask):
    _chk(z, torch.float32, "predictor hidden", 3)
    b, t, f = z.shape
    out = torch.empty(b, t, device=z.device, dtype=torch.float32)
    _launch("lfs2_rowdot_mask", _p(z), _p(w), _p(bias), _p(mask), _p(out), b * t, f, _s(), nbytes=4.0 * b * t * (f + 1))
    return out


def bucket_embed_add_(x, val, std, mean, bins, emb, idx_forced=None, acc=None, acc_init=False, want_idx=False):
    _chk(x, torch.float32, "x", 3)
    b, t, d = x.shape
    if idx_forced is not None:
        _chk(idx_forced, torch.int64, "forced bucket indices")
    else:
        _chk(val, torch.float32, "variance values")
    idx_out = torch.empty(b, t, device=x.device, dtype=torch.int64) if want_idx else None
    mode = 0 if acc is None else (1 if acc_init else 2)
    _launch("lfs2_bucket_embed_add", _p(x), _p(val), float(std), float(mean), _p(bins),
                                                emb.shape[0], _p(emb), _p(idx_forced), _p(idx_out), _p(acc), mode,
                                                b * t, d, _s(), nbytes=4.0 * b * t * d * (2 + (acc is not None)))
    drop_planes(x)
    return idx_out


def decoder_input_planes(x, pe, spk, bucket=None, want_f16=False):
    """((x + emb[bucket(val)]) + pe) + spk as Planes (no fp32 result): lfs2_decoder_input_planes.
    bucket = None or dict(val, std, mean, bins, emb, idx_forced=None, acc=None, acc_init=False, want_idx=False) -- the
    arguments of bucket_embed_add_ for the last frame-level variance encoder.  -> (Planes, bucket indices or None)"""
    _chk(x, torch.float32, "x", 3)
    b, t, d = x.shape
    if t > pe.shape[-2]:
        raise ValueError(f"sequence length {t} exceeds the positional table ({pe.shape[-2]})")
    out = _empty_planes(x.shape, x.device)
    if want_f16:
        out.h = torch.empty(x.shape, device=x.device, dtype=torch.float16)
    k = bucket or {}
    forced, acc = k.get("idx_forced"), k.get("acc")
    if bucket is not None:
        if forced is not None:
            _chk(forced, torch.int64, "forced bucket indices")
        else:
            _chk(k["val"], torch.float32, "variance values")
    idx_out = torch.empty(b, t, device=x.device, dtype=torch.int64) if k.get("want_idx") else None
    mode = 0 if acc is None else (1 if k.get("acc_init") else 2)
    emb = k.get("emb")
    _launch("lfs2_decoder_input_planes", _p(x), _p(k.get("val")), float(k.get("std", 1.0)), float(k.get("mean", 0.0)),
            _p(k.get("bins")), emb.shape[0] if emb is not None else 0, _p(emb), _p(forced), _p(idx_out), _p(acc), mode,
            _p(pe), _p(spk), b, t, d, _p(out.hi), _p(out.lo), _p(out.h), _s(), tag="lfs2_decoder_input_planes",
            nbytes=b * t * d * (4.0 + 4.0 + (2.0 if want_f16 else 0.0) + (4.0 if acc is not None else 0.0)))
    return out, idx_out


# ---- stochastic duration predictor, inference direction (csrc/sdp.cu) ----
def sdp_dwconv(x, pad_mask, wt, bias, dilation):
    """dilated depthwise conv over T on (B,T,C) fp32, PAD rows read as zeros; wt (k, C)"""
    _chk(x, torch.float32, "sdp_dwconv input", 3); _chk(wt, torch.float32, "sdp_dwconv weight", 2)
    b, t, c = x.shape
    out = torch.empty_like(x)
    _launch("lfs2_sdp_dwconv", _p(x), _p(pad_mask), _p(wt), _p(bias), _p(out), b, t, c, wt.shape[0], int(dilation), _s(),
            nbytes=8.0 * x.numel())
    return out


def sdp_ln_gelu(y, gamma, beta, eps=1e-5, res=None):
    """[res +] gelu(LayerNorm over the last dim)"""
    _chk(y, torch.float32, "sdp_ln_gelu input")
    c = y.shape[-1]
    out = torch.empty_like(y)
    _launch("lfs2_sdp_ln_gelu", _p(y), _p(gamma), _p(beta), float(eps), _p(res), _p(out), y.numel() // c, c, _s(),
            nbytes=(12.0 if res is not None else 8.0) * y.numel())
    return out


def sdp_flow_pre(z, channel, w, bias, g):
    """h = z[..., channel, None] * w + bias + g: z (B,T,2), w / bias (C), g (B,T,C)"""
    _chk(z, torch.float32, "flow state", 3); _chk(g, torch.float32, "flow conditioning", 3)
    out = torch.empty_like(g)
    _launch("lfs2_sdp_flow_pre", _p(z), int(channel), _p(w), _p(bias), _p(g), _p(out), g.numel() // g.shape[-1], g.shape[-1],
            _s(), nbytes=8.0 * g.numel())
    return out


def sdp_spline_inverse_(z, x1_channel, h, pad_mask, hidden_channels, tail_bound=5.0):
    """in place on z (B,T,2): channel x1_channel through the inverse spline parameterised by h (B,T,>=29)"""
    _chk(z, torch.float32, "flow state", 3); _chk(h, torch.float32, "spline parameters", 3)
    _launch("lfs2_sdp_spline_inverse", _p(z), int(x1_channel), _p(h), h.shape[-1], _p(pad_mask), int(hidden_channels),
            float(tail_bound), z.numel() // 2, _s())
    return z


def sdp_affine_reverse_(z, translation, log_scale, pad_mask, flip):
    _chk(z, torch.float32, "flow state", 3)
    _launch("lfs2_sdp_affine_reverse", _p(z), _p(translation), _p(log_scale), _p(pad_mask), int(flip), z.numel() // 2, _s())
    return z


def sdp_durations(logw, src_mask):
    _chk(logw, torch.float32, "log-durations", 2); _chk(src_mask, torch.bool, "src_mask", 2)
    dur = torch.empty(logw.shape, device=logw.device, dtype=torch.int32)
    _launch("lfs2_sdp_durations", _p(logw), _p(src_mask), _p(dur), logw.shape[0], logw.shape[1], _s())
    return dur


def prior_embed(prior, bins, emb):
    """PriorEmbedding: prior (B) fp32 -> (relu(emb[bucketize(prior)]) (B,d), bucket indices (B) int64)"""
    _chk(prior, torch.float32, "prior values", 1)
    b, d = prior.shape[0], emb.shape[1]
    out = torch.empty(b, d, device=prior.device, dtype=torch.float32)
    idx = torch.empty(b, device=prior.device, dtype=torch.int64)
    _launch("lfs2_prior_embed", _p(prior), _p(bins), emb.shape[0], _p(emb), _p(out), _p(idx), b, d, _s())
    return out, idx


def duration_round_guard(log_dur, src_mask):
    _chk(log_dur, torch.float32, "duration_prediction", 2); _chk(src_mask, torch.bool, "src_mask", 2)
    dur = torch.empty(log_dur.shape, device=log_dur.device, dtype=torch.int32)
    _launch("lfs2_duration_round_guard", _p(log_dur), _p(src_mask), _p(dur), log_dur.shape[0],
                                                    log_dur.shape[1], _s())
    return dur


def length_regulate_scan(durations, batch_first_shape):
    """prefix sums of the durations: -> (cum (B,Tp) int64, lengths (B) int64, max_len (1) int64, all on the device)"""
    if durations.dtype not in (torch.int32, torch.int64):
        raise TypeError("length_regulate: durations must be int32 or int64")
    durations = durations.contiguous()
    b, tp = batch_first_shape
    if tuple(durations.shape) != (b, tp):
        raise ValueError(f"length_regulate: durations {tuple(durations.shape)} do not match x[:2] = {(b, tp)}")
    dev = durations.device
    cum = torch.empty(b, tp, device=dev, dtype=torch.int64)
    lengths = torch.empty(b, device=dev, dtype=torch.int64)
    mx = torch.empty(1, device=dev, dtype=torch.int64)
    _launch("lfs2_length_regulate_scan", _p(durations), int(durations.dtype == torch.int64), _p(cum), _p(lengths),
            _p(mx), b, tp, _s())
    return cum, lengths, mx


def pack_valid_rows(x, lengths, total):
    """x (B,L,W) fp32, lengths (B) int64 on the device, total = sum(min(lengths, L)) (known on the host) ->
    (total, W): every utterance's valid rows back to back"""
    _chk(x, torch.float32, "pack_valid_rows input", 3); _chk(lengths, torch.int64, "frame counts", 1)
    b, l, w = x.shape
    out = torch.empty(int(total), w, device=x.device, dtype=torch.float32)
    _launch("lfs2_pack_valid_rows", _p(x), _p(lengths), _p(out), b, l, w, _s(), nbytes=8.0 * int(total) * w)
    return out


def length_regulate_scatter(x, cum, lengths, l, cap):
    """out (B,l,d), mask (B,l): frames below min(lengths[b], cap) copy their phone's row, the rest are PAD (+0)"""
    b, tp, d = x.shape
    out = torch.empty(b, l, d, device=x.device, dtype=x.dtype)
    mask = torch.empty(b, l, device=x.device, dtype=torch.bool)
    _launch("lfs2_length_regulate_scatter_ex", _p(x), _p(cum), _p(lengths), _p(out), _p(mask), b, tp, l, cap,
            d * x.element_size(), _s(), tag="lfs2_length_regulate_scatter",
            nbytes=float(b * tp * (d * x.element_size() + 8) + b * l * (d * x.element_size() + 1)))
    return out, mask


def length_regulate(x, durations, max_length, scan=None, frames=None, pad_to_multiple_of=None):
    """LengthRegulator.forward: x (B,Tp,d) any dtype, durations (B,Tp) int32/int64 ->
    (out (B,L,d), mask (B,L) bool), L = min(longest, int(max_length)).  One host read-back (the maximum length).
    scan = a precomputed length_regulate_scan result; frames = (l, cap) overrides the output length
    (length-bucketed synthesis: l >= cap, frames in [cap, l) are PAD) and needs no read-back."""
    if not (x.is_cuda and x.is_contiguous() and x.dim() == 3):
        raise _lib.Lfs2Error("length_regulate: x must be a contiguous CUDA (B,Tp,d) tensor")
    cum, lengths, mx = scan if scan is not None else length_regulate_scan(durations, x.shape[:2])
    if frames is None:
        longest = int(mx.item())  # the single device->host sync of the path
        l = cap = min(longest, int(max_length)) if max_length is not None else longest
        if pad_to_multiple_of is not None:
            l = cap = -(-l // pad_to_multiple_of) * pad_to_multiple_of
    else:
        l, cap = frames
    return length_regulate_scatter(x, cum, lengths, l, cap)


# ---------------------------------------------------------------------------------------------
# tensor-core path: bf16 hi/lo planes + tcgen05 GEMM with fused epilogues
class Planes:
    """bf16 hi/lo planes of an fp32 tensor (x = hi + lo): the operand format of lfs2_gemm_tc.
    A single 16-bit plane travels as Planes(hi = fp16 / bf16 tensor, lo = None).  `h` (optional): the same values as ONE
    fp16 plane next to the hi/lo pair -- the activation operand of the 2-pass GEMM recipe (npass = 2)."""

    __slots__ = ("hi", "lo", "h")

    def __init__(self, hi, lo, h=None):
        self.hi, self.lo, self.h = hi, lo, h

    @property
    def shape(self):
        return self.hi.shape

    def float(self):  # for tests
        return self.hi.float() + self.lo.float()


def _empty_planes(shape, dev):
    """hi and lo planes carved out of ONE allocation (one caching-allocator call per launch instead of two)"""
    buf = torch.empty((2,) + tuple(shape), device=dev, dtype=torch.bfloat16)
    return Planes(buf[0], buf[1])


def attach_planes(x, planes):
    """Remember the hi/lo planes a kernel already produced for the fp32 tensor x, so the next
    tensor-core GEMM does not have to split it again (module APIs stay tensor -> tensor)."""
    if planes is not None:
        x._lfs2_planes = planes
    return x


def drop_planes(x):
    if getattr(x, "_lfs2_planes", None) is not None:
        x._lfs2_planes = None


def planes_of(x, want_f16=False):
    """the Planes of fp32 tensor x (remembered ones if a kernel already produced them); want_f16: with the fp16 plane
    (Planes.h) the 2-pass GEMM recipe reads"""
    p = getattr(x, "_lfs2_planes", None)
    if p is not None and (not want_f16 or p.h is not None):
        return p
    return split_bf16(x.contiguous(), want_f16=want_f16)


def _limited_fraction(row_limit, t):
    """fraction of the (B, t) rows a row-limited launch really processes (profiling only: reads the lengths back)"""
    if row_limit is None or PROFILE is None:
        return 1.0
    lim, extra = row_limit[0], row_limit[1]
    rows = torch.clamp((lim.long() + extra + 127) // 128 * 128, max=t).clamp(min=0).sum().item()
    return rows / float(lim.numel() * t)


def mask_lengths(mask):
    """(B,T) bool padding mask (True = PAD) -> int32 (B): 1 + index of the last non-PAD position"""
    _chk(mask, torch.bool, "padding mask", 2)
    out = torch.empty(mask.shape[0], device=mask.device, dtype=torch.int32)
    _launch("lfs2_mask_lengths", _p(mask), _p(out), mask.shape[0], mask.shape[1], _s())
    return out


def zero_masked_rows_(x, mask):
    """x (..., w) fp32, mask (...) bool: x[mask] = 0 in place"""
    _chk(x, torch.float32, "zero_masked_rows input"); _chk(mask, torch.bool, "zero_masked_rows mask")
    if tuple(x.shape[:-1]) != tuple(mask.shape):
        raise ValueError("zero_masked_rows_: mask must be shaped like x without its last dim")
    _launch("lfs2_zero_masked_rows", _p(x), _p(mask), mask.numel(), x.shape[-1], _s(), nbytes=2.0 * x.numel())
    return x


def dwconv1d_planes(x, wt, bias, out="planes", row_limit=None):
    """depthwise conv, x (B,T,d) fp32 tensor or Planes, wt (ksize,d) -> Planes (fp32 if out == "f32"; ONE fp16 plane,
    Planes(hi = fp16, lo = None), if out == "f16").
    row_limit = (lengths int32 (B), extra): 128-row groups starting at or after lengths[b] + extra are skipped."""
    _chk(wt, torch.float32, "dwconv weight", 2)
    if isinstance(x, Planes):
        _chk(x.hi, torch.bfloat16, "dwconv input", 3); _chk(x.lo, torch.bfloat16, "dwconv input", 3)
        xf, xh, xl, shape, dev = None, x.hi, x.lo, x.hi.shape, x.hi.device
    else:
        _chk(x, torch.float32, "dwconv input", 3)
        xf, xh, xl, shape, dev = x, None, None, x.shape, x.device
    b, t, d = shape
    of = torch.empty(shape, device=dev, dtype=torch.float32) if out == "f32" else None
    po = _empty_planes(shape, dev) if out == "planes" else None
    ph = torch.empty(shape, device=dev, dtype=torch.float16) if out == "f16" else None  # ONE fp16 plane (2-pass GEMMs)
    lim, extra = (row_limit[0], row_limit[1]) if row_limit is not None else (None, 0)
    frac = _limited_fraction(row_limit, t)
    _launch("lfs2_dwconv1d_planes_ex", _p(xf), _p(xh), _p(xl), _p(wt), _p(bias), _p(of), _p(po.hi if po else None),
            _p(po.lo if po else None), _p(ph), b, t, d, wt.shape[0], _p(lim), int(extra), _s(), tag="lfs2_dwconv1d",
            flops=2.0 * b * t * d * wt.shape[0] * frac, nbytes=(6.0 if out == "f16" else 8.0) * b * t * d * frac)
    return of if out == "f32" else (Planes(ph, None) if out == "f16" else po)


def merge_planes(p):
    """Planes -> fp32 tensor (hi + lo)"""
    _chk(p.hi, torch.bfloat16, "planes.hi"); _chk(p.lo, torch.bfloat16, "planes.lo")
    out = torch.empty(p.hi.shape, device=p.hi.device, dtype=torch.float32)
    _launch("lfs2_merge_planes", _p(p.hi), _p(p.lo), _p(out), p.hi.numel(), _s(), nbytes=8.0 * p.hi.numel())
    return attach_planes(out, p)


def split_bf16(x, want_f16=False):
    _chk(x, torch.float32, "split_bf16 input")
    pl = _empty_planes(x.shape, x.device)
    if want_f16:
        pl.h = torch.empty(x.shape, device=x.device, dtype=torch.float16)
    _launch("lfs2_split_bf16_ex", _p(x), _p(pl.hi), _p(pl.lo), _p(pl.h), x.numel(), _s(), tag="lfs2_split_bf16",
            nbytes=(10.0 if want_f16 else 8.0) * x.numel())
    return pl


def split_f16(x):
    """fp32 -> fp16 hi/lo planes (hi + lo = x to ~2^-22): the weight operand of the 2-pass recipe (npass = 2)"""
    _chk(x, torch.float32, "split_f16 input")
    buf = torch.empty((2,) + tuple(x.shape), device=x.device, dtype=torch.float16)
    _launch("lfs2_split_f16", _p(x), _p(buf[0]), _p(buf[1]), x.numel(), _s(), nbytes=8.0 * x.numel())
    return Planes(buf[0], buf[1])


_IDENT = {}


def _identity_planes(n, device):
    key = (n, str(device))
    if key not in _IDENT:
        _IDENT[key] = torch.eye(n, device=device, dtype=torch.bfloat16).contiguous()
    return _IDENT[key]


OUT_KINDS = {"planes": 0, "f32": 1, "f16": 2, "bf16": 3}


def gemm_tc(a, w, bias, taps=1, relu=False, residual=None, gamma=None, beta=None, eps=LN_EPS, out="f32",
            npass=3, tag=None, row_limit=None, dilation=1, leaky_slope=None, row_mask=None, zero_skipped=True):
    """a: Planes (B,T,d) [taps>1: Conv1d over T per utterance] or (...,d) for taps == 1;
    w: Planes (n, taps*d); residual: Planes shaped like the output (added on the tensor core).
    out = "f32" -> fp32 tensor, "planes" -> Planes (bf16 hi/lo), "f16" -> Planes(hi = ONE fp16 tensor, lo = None: the
    operand of the single-pass fp16 attention), "bf16" -> Planes(hi = ONE bf16 tensor, lo = None: a result that only
    feeds npass = 1 products); shaped like a with last dim n.
    dilation: tap spacing (dilated Conv1d); leaky_slope: leaky ReLU instead of ReLU; row_mask (B,T) bool: rows written
    as zeros (PAD frames of a ragged batch).
    npass = 3: hi.hi + lo.hi + hi.lo; 1: hi.hi; 2: a as ONE fp16 plane against fp16 hi/lo weight planes (split_f16),
    a.w_hi + a.w_lo (no LayerNorm / residual epilogue in this recipe)."""
    if not isinstance(a, Planes) or not isinstance(w, Planes):
        raise TypeError("gemm_tc: operands must be Planes (see split_bf16)")
    if out not in OUT_KINDS:
        raise ValueError("gemm_tc: out must be 'f32', 'planes', 'f16' or 'bf16'")
    d = a.shape[-1]
    n = w.shape[0]
    if w.shape[1] != taps * d:
        raise ValueError(f"gemm_tc: weight {tuple(w.shape)} does not match taps*d = {taps * d}")
    if taps == 1 and row_limit is None and row_mask is None:
        batch, t = 1, a.hi.numel() // d
    else:
        if a.hi.dim() != 3:
            raise ValueError("gemm_tc: conv mode needs a (B,T,d) operand")
        batch, t = a.shape[0], a.shape[1]
    if npass == 2:  # the activation operand of the 2-pass recipe: ONE fp16 plane (a Planes' .h, or Planes(fp16, None))
        if a.hi.dtype != torch.float16:
            if a.h is None:
                raise ValueError("gemm_tc: npass = 2 needs the fp16 plane of a (Planes.h or Planes(fp16, None))")
            a = Planes(a.h, None)
        _chk(a.hi, torch.float16, "gemm_tc: a of the 2-pass recipe")
        _chk(w.hi, torch.float16, "gemm_tc: w of the 2-pass recipe (split_f16)")
        _chk(w.lo, torch.float16, "gemm_tc: w of the 2-pass recipe (split_f16)")
    else:
        for x_ in (a.hi, a.lo, w.hi, w.lo):
            if x_ is not None or npass == 3:  # (single-plane operands: Planes(hi, None) are fine for npass = 1)
                _chk(x_, torch.bfloat16, "gemm_tc operand plane")
    out_shape = tuple(a.shape[:-1]) + (n,)
    dev = a.hi.device
    # an fp32 result of a row-limited launch is a user-visible tensor: the rows of skipped tiles read as zeros
    # (zero_skipped=False: the consumer skips the same rows -- the row-limited LayerNorm of the wide FFTBlock)
    of = (torch.zeros if row_limit is not None and zero_skipped else torch.empty)(out_shape, device=dev,
                                                                                  dtype=torch.float32) \
        if out == "f32" else None
    po = _empty_planes(out_shape, dev) if out == "planes" else None
    if out in ("f16", "bf16"):
        po = Planes(torch.empty(out_shape, device=dev, dtype=torch.float16 if out == "f16" else torch.bfloat16), None)
    ident = None
    if residual is not None:
        if not isinstance(residual, Planes) or tuple(residual.shape) != out_shape:
            raise ValueError("gemm_tc: residual must be Planes shaped like the output")
        _chk(residual.hi, torch.bfloat16, "gemm_tc residual"); _chk(residual.lo, torch.bfloat16, "gemm_tc residual")
        ident = _identity_planes(n, dev)
    if row_mask is not None:
        _chk(row_mask, torch.bool, "gemm_tc row mask", 2)
        if tuple(row_mask.shape) != (batch, t):
            raise ValueError("gemm_tc: row_mask must be (B, T)")
    m = batch * t
    lim, extra = (row_limit[0], row_limit[1]) if row_limit is not None else (None, 0)
    m = m * _limited_fraction(row_limit, t)  # rows really processed (for the flop / byte accounting below)
    ws = None
    if lim is not None:
        # row_limit may carry a dict as third element: the tile list built by the first launch is reused by the next
        cache = row_limit[2] if len(row_limit) > 2 else None
        key = (batch, t)
        if cache is not None and key in cache:
            ws, lim = cache[key], None
        else:
            ws = torch.empty(_lib.lib().lfs2_gemm_tc_limited_workspace_bytes(batch, t) // 4, device=dev, dtype=torch.int32)
            if cache is not None:
                cache[key] = ws
    act = 0 if not (relu or leaky_slope is not None) else (2 if leaky_slope is not None else 1)
    out_bytes = {"f32": 4.0, "planes": 4.0, "f16": 2.0, "bf16": 2.0}[out]
    _launch("lfs2_gemm_tc_ex", _p(a.hi), _p(a.lo), batch, t, d, taps, int(dilation), _p(w.hi), _p(w.lo), n, _p(bias), act,
            float(leaky_slope or 0.0), _p(residual.hi if residual is not None else None),
            _p(residual.lo if residual is not None else None), _p(ident), _p(gamma), _p(beta), float(eps),
            _p(of if of is not None else po.hi), _p(po.lo if out == "planes" else None), OUT_KINDS[out], npass, _p(lim),
            int(extra), _p(ws), _p(row_mask), _s(), tag=tag or f"gemm_tc_n{n}_k{taps * d}",
            flops=2.0 * m * n * taps * d, passes=npass,
            nbytes=(2.0 if npass == 2 else 4.0) * m * d + 4.0 * n * taps * d + out_bytes * m * n
            + (4.0 * m * n if residual is not None else 0.0))
    return of if out == "f32" else po


def predictor_layer_tc(a, w, bias, gamma, beta, eps=LN_EPS, npass=3, next_dw=None, head=None, row_limit=None):
    """One depthwise predictor layer with its successor fused into the LayerNorm epilogue (lfs2_predictor_layer_tc):
    a Planes (B,T,256) = depthwise3 of the previous layer, w Planes (256,256).
    next_dw = (wt (3,256), bias (256)) -> Planes: depthwise3 of the NEXT layer applied to LayerNorm(relu(a.w^T + bias));
    head = (w (1,256) or (256), b (1), mask (B,T) bool or None) -> (B,T) fp32: the Linear(256,1) head, masked."""
    if (next_dw is None) == (head is None):
        raise ValueError("predictor_layer_tc: exactly one of next_dw / head")
    if a.hi.dim() != 3 or a.shape[-1] != 256 or tuple(w.shape) != (256, 256):
        raise ValueError("predictor_layer_tc: needs (B,T,256) activations and a (256,256) weight")
    for x_ in (a.hi, a.lo, w.hi, w.lo):
        _chk(x_, torch.bfloat16, "predictor_layer_tc operand plane")
    b, t, d = a.shape
    dev = a.hi.device
    lim, extra = (row_limit[0], row_limit[1]) if row_limit is not None else (None, 0)
    ws = None
    if lim is not None:
        cache = row_limit[2] if len(row_limit) > 2 else None
        key = ("pl", b, t, next_dw is not None)   # the stencil form walks 126-row tiles: its own list
        if cache is not None and key in cache:
            ws, lim = cache[key], None
        else:
            ws = torch.empty(_lib.lib().lfs2_predictor_layer_tc_workspace_bytes(b, t) // 4, device=dev, dtype=torch.int32)
            if cache is not None:
                cache[key] = ws
    m = b * t * _limited_fraction(row_limit, t)
    po = hout = None
    dw_w = dw_b = hw = hb = hmask = None
    if next_dw is not None:
        dw_w, dw_b = next_dw
        _chk(dw_w, torch.float32, "next depthwise weight", 2)
        if tuple(dw_w.shape) != (3, 256):
            raise ValueError("predictor_layer_tc: the fused depthwise conv has kernel size 3")
        po = _empty_planes((b, t, 256), dev)
    else:
        hw, hb, hmask = head
        if hmask is not None:
            _chk(hmask, torch.bool, "head mask", 2)
        # rows of skipped tiles are PAD positions: the head masks them to 0
        hout = (torch.zeros if row_limit is not None else torch.empty)(b, t, device=dev, dtype=torch.float32)
    _launch("lfs2_predictor_layer_tc", _p(a.hi), _p(a.lo), b, t, _p(w.hi), _p(w.lo), _p(bias), _p(gamma), _p(beta),
            float(eps), npass, _p(dw_w), _p(dw_b), _p(po.hi if po else None), _p(po.lo if po else None), _p(hw), _p(hb),
            _p(hmask), _p(hout), _p(lim), int(extra), _p(ws), _s(), tag="predictor_pw_ln_gemm", flops=2.0 * m * 256 * 256,
            passes=npass, nbytes=4.0 * m * 256 + 4.0 * 256 * 256 + (4.0 * m * 256 if po is not None else 4.0 * m))
    return po if po is not None else hout


def ffn_fused_tc(u, w1, b1, w2, b2, residual, gamma, beta, eps=LN_EPS, npass=3, row_limit=None, want_f16=False):
    """LayerNorm(residual + relu(u . w1^T + b1) . w2^T + b2) in one kernel, the F-wide intermediate on chip.
    u, residual: Planes (..., 256); w1 Planes (F, 256); w2 Planes (256, F) -> Planes (..., 256).
    npass = 2: u is ONE fp16 plane (Planes(hi = fp16, lo = None)), w1 / w2 fp16 hi/lo planes (split_f16), products
    a.w_hi + a.w_lo.  want_f16: the result also
    carries its fp16 plane (Planes.h) for the next block's 2-pass QKV GEMM.
    row_limit = (lengths int32 (B), extra[, cache dict]) on (B,T,256) operands: 128-row tiles that hold no row
    t < roundup128(lengths[b] + extra) of any utterance are skipped (their output rows stay unwritten)."""
    d = u.shape[-1]
    f = w1.shape[0]
    if d != 256 or tuple(w1.shape) != (f, 256) or tuple(w2.shape) != (256, f) or tuple(residual.shape) != tuple(u.shape):
        raise ValueError("ffn_fused_tc: shapes must be u/residual (..., 256), w1 (F, 256), w2 (256, F)")
    if npass == 2:
        _chk(u.hi, torch.float16, "ffn_fused_tc: u of the 2-pass recipe (one fp16 plane)")
    else:
        _chk(u.hi, torch.bfloat16, "ffn_fused_tc operand plane")
        if npass == 3:
            _chk(u.lo, torch.bfloat16, "ffn_fused_tc operand plane")
    for x_ in (w1.hi, w1.lo, w2.hi, w2.lo):
        _chk(x_, torch.float16 if npass == 2 else torch.bfloat16, "ffn_fused_tc weight plane")
    for x_ in (residual.hi, residual.lo):
        _chk(x_, torch.bfloat16, "ffn_fused_tc residual plane")
    m = u.hi.numel() // d
    out = _empty_planes(tuple(u.shape), u.hi.device)
    if want_f16:
        out.h = torch.empty(tuple(u.shape), device=u.hi.device, dtype=torch.float16)
    ident = _identity_planes(d, u.hi.device)
    if row_limit is None:
        batch, t, lim, extra, ws = 1, m, None, 0, None
    else:
        if u.hi.dim() != 3:
            raise ValueError("ffn_fused_tc: a row limit needs (B,T,256) operands")
        batch, t = u.shape[0], u.shape[1]
        lim, extra = row_limit[0], row_limit[1]
        cache = row_limit[2] if len(row_limit) > 2 else None
        key = ("ffn", batch, t)
        if cache is not None and key in cache:
            ws, lim = cache[key], None
        else:
            ws = torch.empty(_lib.lib().lfs2_ffn_fused_tc_limited_workspace_bytes(batch, t) // 4, device=u.hi.device,
                             dtype=torch.int32)
            if cache is not None:
                cache[key] = ws
        m = m * _limited_fraction(row_limit, t)
    _launch("lfs2_ffn_fused_tc_ex", _p(u.hi), _p(u.lo if npass == 3 else None), batch, t, _p(w1.hi), _p(w1.lo), f, _p(b1),
            _p(w2.hi), _p(w2.lo), _p(b2), _p(residual.hi), _p(residual.lo), _p(ident), _p(gamma), _p(beta), float(eps),
            _p(out.hi), _p(out.lo), _p(out.h), npass, _p(lim), int(extra), _p(ws), _s(), tag="ffn_fused",
            flops=4.0 * m * d * f, passes=npass,
            nbytes=m * d * ((2.0 if npass == 2 else 4.0) + 8.0 + (2.0 if want_f16 else 0.0)) + 8.0 * d * f)
    return out


def attention_tc(qkv, kpm, nhead, npass=3, want_f32=False, want_planes=True, row_limit=None):
    """qkv: Planes (B,T,3d) packed [q|k|v] -- bf16 hi/lo planes, or ONE fp16 plane (Planes(hi = fp16 tensor, lo = None)
    from gemm_tc(out="f16"): single-pass fp16 products, npass is ignored); kpm (B,T) bool True=PAD ->
    (ctx f32 or None, ctx Planes or None).
    row_limit = (lengths int32 (B), extra, ...): 128-row query tiles starting at or after lengths[b] + extra are
    skipped (their ctx rows stay unwritten)."""
    if not isinstance(qkv, Planes):
        raise TypeError("attention_tc: qkv must be Planes")
    f16 = qkv.hi.dtype == torch.float16
    if f16:
        _chk(qkv.hi, torch.float16, "qkv (fp16 plane)", 3)
        npass = 1
    else:
        _chk(qkv.hi, torch.bfloat16, "qkv.hi", 3); _chk(qkv.lo, torch.bfloat16, "qkv.lo", 3)
    b, t, d3 = qkv.shape
    d = d3 // 3
    if kpm is not None:
        _chk(kpm, torch.bool, "key_padding_mask", 2)
    dev = qkv.hi.device
    ctx = torch.empty(b, t, d, device=dev, dtype=torch.float32) if want_f32 else None
    po = _empty_planes((b, t, d), dev) if want_planes else None
    ws = torch.empty(max(1, _lib.lib().lfs2_attention_tc_workspace_bytes(b)), device=dev, dtype=torch.uint8)
    lim, extra = (row_limit[0], row_limit[1]) if row_limit is not None else (None, 0)
    fl, frac = 0.0, 1.0
    if PROFILE is not None:  # algorithmic flops: every (computed) query row x the utterance's VALID keys
        nkeys = (~kpm).sum(1).double() if kpm is not None else torch.full((b,), float(t), device=dev, dtype=torch.float64)
        rows = torch.full((b,), float(t), device=dev, dtype=torch.float64)
        if lim is not None:
            rows = torch.clamp((lim.long() + extra + 127) // 128 * 128, min=0, max=t).double()
            frac = float(rows.sum()) / float(b * t)
        fl = float(4.0 * d * (rows * nkeys.to(dev)).sum())
    in_bytes = (2.0 if f16 else 4.0) * qkv.hi.numel()
    _launch("lfs2_attention_tc_ex", _p(qkv.hi), _p(qkv.lo), 1 if f16 else 0, _p(kpm), _p(po.hi if po else None),
            _p(po.lo if po else None), _p(ctx), _p(ws), b, t, d, nhead, npass, _p(lim), int(extra), _s(),
            tag="lfs2_attention_tc", flops=fl, passes=npass,
            nbytes=(in_bytes + 4.0 * b * t * d * (int(want_f32) + int(want_planes))) * frac)
    return ctx, po


def attention_tc_wide(qkv, kpm, nhead, want_f32=False, want_planes=True, row_limit=None):
    """flash attention for head_dim 256 / 384.  qkv: ONE 16-bit tensor (B,T,3d) [q|k|v], torch.bfloat16 or torch.float16
    (or Planes: the hi plane is used) -> (ctx f32 or None, ctx Planes or None)"""
    if isinstance(qkv, Planes):
        qkv = qkv.hi
    if qkv.dtype not in (torch.bfloat16, torch.float16):
        raise TypeError("attention_tc_wide: qkv must be a bf16 or fp16 tensor")
    _chk(qkv, qkv.dtype, "qkv", 3)
    b, t, d3 = qkv.shape
    d = d3 // 3
    if kpm is not None:
        _chk(kpm, torch.bool, "key_padding_mask", 2)
    dev = qkv.device
    ctx = torch.empty(b, t, d, device=dev, dtype=torch.float32) if want_f32 else None
    po = _empty_planes((b, t, d), dev) if want_planes else None
    ws = torch.empty(max(1, _lib.lib().lfs2_attention_tc_workspace_bytes(b)), device=dev, dtype=torch.uint8)
    lim, extra = (row_limit[0], row_limit[1]) if row_limit is not None else (None, 0)
    fl = 0.0
    if PROFILE is not None:
        nkeys = (~kpm).sum(1).double() if kpm is not None else torch.full((b,), float(t), device=dev, dtype=torch.float64)
        fl = float(4.0 * d * t * nkeys.sum())
    _launch("lfs2_attention_tc_wide", _p(qkv), 1 if qkv.dtype == torch.float16 else 0, _p(kpm), _p(po.hi if po else None),
            _p(po.lo if po else None), _p(ctx), _p(ws), b, t, d, nhead, _p(lim), int(extra), _s(),
            tag="lfs2_attention_tc", flops=fl, nbytes=2.0 * qkv.numel() + 4.0 * b * t * d * (int(want_f32) + int(want_planes)))
    return ctx, po


# ---------------------------------------------------------------------------------------------
# FastDiff variance adaptor glue
def diffusion_step_embed(steps, dim):
    """steps (B) fp32 -> (B, dim) sin / cos embedding"""
    _chk(steps, torch.float32, "diffusion steps", 1)
    out = torch.empty(steps.shape[0], dim, device=steps.device, dtype=torch.float32)
    _launch("lfs2_diffusion_step_embed", _p(steps), _p(out), steps.shape[0], dim, _s())
    return out


def swish_(x):
    _chk(x, torch.float32, "swish input")
    _launch("lfs2_swish", _p(x), x.numel(), _s())
    return x


def diffusion_input(xt, w_in, b_in, c, noise_embed):
    """xt (B,T), w_in / b_in (d), c (B,T,d), noise_embed (B,d) -> (B,T,d)"""
    _chk(xt, torch.float32, "noisy track", 2); _chk(c, torch.float32, "condition", 3)
    b, t, d = c.shape
    out = torch.empty_like(c)
    _launch("lfs2_diffusion_input", _p(xt), _p(w_in), _p(b_in), _p(c), _p(noise_embed), _p(out), b, t, d, _s(),
            nbytes=4.0 * b * t * (2 * d + 1))
    return out


def diffusion_mix(x, a=None, y=None, e=None, s=None, z=None, g=None, add=0.0, zero_mask=None):
    """(a[b] x + e[b] y) * s[b] + g[b] z + add on (B,T) tracks with (B) coefficient vectors; zero where zero_mask"""
    _chk(x, torch.float32, "track", 2)
    for v in (y, z):
        if v is not None:
            _chk(v, torch.float32, "track", 2)
    for v in (a, e, s, g):
        if v is not None:
            _chk(v, torch.float32, "per-utterance coefficients", 1)
    if zero_mask is not None:
        _chk(zero_mask, torch.bool, "mask", 2)
    out = torch.empty_like(x)
    _launch("lfs2_diffusion_mix", _p(x), _p(y), _p(z), _p(a), _p(e), _p(s), _p(g), float(add), _p(zero_mask), _p(out),
            x.shape[0], x.shape[1], _s())
    return out


# ---------------------------------------------------------------------------------------------
# HiFi-GAN generator glue (the convolutions are gemm_tc launches)
def lrelu_planes(x, slope):
    out = _empty_planes(tuple(x.shape), x.hi.device)
    _launch("lfs2_lrelu_planes", _p(x.hi), _p(x.lo), _p(out.hi), _p(out.lo), x.hi.numel(), float(slope), _s(),
            nbytes=8.0 * x.hi.numel())
    return out


def mean3_lrelu_planes(a, b, c, slope):
    """leaky_relu((a + b + c) / 3, slope) on Planes"""
    out = _empty_planes(tuple(a.shape), a.hi.device)
    _launch("lfs2_mean3_lrelu_planes", _p(a.hi), _p(a.lo), _p(b.hi), _p(b.lo), _p(c.hi), _p(c.lo), _p(out.hi), _p(out.lo),
            a.hi.numel(), 1.0 / 3.0, float(slope), _s(), nbytes=16.0 * a.hi.numel())
    return out


def mel_to_planes(mel, lengths, c_padded):
    """mel (B, C, T) fp32 channels-first -> Planes (B, T, c_padded), zero beyond lengths (int32 (B) or None)"""
    _chk(mel, torch.float32, "mel", 3)
    b, c, t = mel.shape
    out = _empty_planes((b, t, c_padded), mel.device)
    _launch("lfs2_mel_to_planes", _p(mel), _p(lengths), _p(out.hi), _p(out.lo), b, c, t, c_padded, _s(),
            nbytes=4.0 * mel.numel() + 4.0 * b * t * c_padded)
    return out


def conv_post_tanh(x, w, bias, lengths, slope):
    """x Planes (B, T, C), w (k*C) tap-major fp32, bias (1) -> tanh(conv(leaky_relu(x))) (B, T) fp32"""
    b, t, c = x.shape
    k = w.numel() // c
    out = torch.empty(b, t, device=x.hi.device, dtype=torch.float32)
    _launch("lfs2_conv_post_tanh", _p(x.hi), _p(x.lo), _p(w), _p(bias), _p(lengths), float(slope), _p(out), b, t, c, k,
            _s(), flops=2.0 * b * t * c * k, nbytes=4.0 * b * t * c + 4.0 * b * t)
    return out


# ---------------------------------------------------------------------------------------------
# train-step config: forward variants that save what the backward needs, backward kernels, loss,
# optimizer.  Parameter gradients accumulate (+=) into the tensors passed as d<param>.
def add_layernorm_train(x, y, gamma, beta, eps=LN_EPS, drop=None):
    """-> (out, z = x (+ y), stats (m,2) = per-row (mean, rstd)); drop = (p, seed, site): z = x + dropout(y), fused"""
    _chk(x, torch.float32, "layernorm input")
    d = x.shape[-1]
    m = x.numel() // d
    out = torch.empty_like(x)
    z = torch.empty_like(x) if y is not None else x
    stats = torch.empty(m, 2, device=x.device, dtype=torch.float32)
    dp, dseed, dsite = drop if drop is not None else (0.0, 0, 0)
    _launch("lfs2_add_layernorm_train", _p(x), _p(y), _p(gamma), _p(beta), _p(out), _p(z if y is not None else None),
            _p(stats), m, d, eps, float(dp), int(dseed), int(dsite), _s(), tag="lfs2_add_layernorm",
            nbytes=4.0 * m * d * (4 if y is not None else 2))
    return out, z, stats


def layernorm_bwd(dy, z, stats, gamma, dgamma, dbeta, add=None, drop=None, dy2=None):
    """-> dz, or (dz, dropout(dz)) when drop = (p, seed, site) of the branch dropout fused into the forward.
    dy2: a second summand of the incoming gradient (a residual join upstream), added on load."""
    _chk(dy, torch.float32, "layernorm_bwd dy"); _chk(z, torch.float32, "layernorm_bwd z")
    if dy2 is not None:
        _chk(dy2, torch.float32, "layernorm_bwd dy2")
        if tuple(dy2.shape) != tuple(dy.shape):
            raise ValueError("layernorm_bwd: dy2 must be shaped like dy")
    d = z.shape[-1]
    m = z.numel() // d
    dz = torch.empty_like(z)
    dzd = torch.empty_like(z) if drop is not None else None
    dp, dseed, dsite = drop if drop is not None else (0.0, 0, 0)
    _launch("lfs2_layernorm_bwd_ex", _p(dy), _p(dy2), _p(z), _p(stats), _p(gamma), _p(add), _p(dz), _p(dzd), _p(dgamma),
            _p(dbeta), m, d, float(dp), int(dseed), int(dsite), _s(), tag="lfs2_layernorm_bwd",
            nbytes=4.0 * m * d * (3 + (add is not None) + (drop is not None) + (dy2 is not None)))
    return dz if drop is None else (dz, dzd)


def gemm_tn_(dw, dy, x, t=0, shift=0, col_offset=0):
    """dw[:, col_offset:col_offset+k] += dy (m,n)^T . x (m,k)   (x rows shifted by `shift` inside utterances of t)"""
    n, k = dy.shape[-1], x.shape[-1]
    m = dy.numel() // n
    _launch("lfs2_gemm_tn", _p(dy), _p(x), ctypes.c_void_p(dw.data_ptr() + 4 * col_offset), m, n, k, n, k,
            dw.shape[-1], t, shift, _s(), tag=f"wgrad_n{n}_k{k}", flops=2.0 * m * n * k, nbytes=4.0 * m * (n + k))


def colsum_(out, a):
    n = a.shape[-1]
    _launch("lfs2_colsum", _p(a), _p(out), a.numel() // n, n, _s(), nbytes=4.0 * a.numel())


def relu_bwd_(dy, y, scale=1.0):
    """dy <- (y > 0 ? dy * scale : 0); scale = 1/(1-p) folds the backward of a dropout applied right after the ReLU"""
    _launch("lfs2_relu_bwd_scaled", _p(dy), _p(y), _p(dy), dy.numel(), float(scale), _s(), tag="lfs2_relu_bwd",
            nbytes=12.0 * dy.numel())
    drop_planes(dy)
    return dy


def relu_bwd_planes(dy, y, scale=1.0, db=None):
    """-> Planes of (y > 0 ? dy * scale : 0); db (cols) += its column sums.  y: the ReLU output as an fp32 tensor or as
    Planes (only the sign of the hi plane is read).  One pass instead of relu_bwd_ + split_bf16 + colsum_."""
    _chk(dy, torch.float32, "relu_bwd_planes dy")
    cols = dy.shape[-1]
    rows = dy.numel() // cols
    if isinstance(y, Planes):
        if y.hi.dtype != torch.bfloat16 or tuple(y.hi.shape) != tuple(dy.shape) or not y.hi.is_contiguous():
            raise ValueError("relu_bwd_planes: y must be contiguous bf16 planes shaped like dy")
        y32, yhi = None, y.hi
    else:
        _chk(y, torch.float32, "relu_bwd_planes y")
        if tuple(y.shape) != tuple(dy.shape):
            raise ValueError("relu_bwd_planes: y must be shaped like dy")
        y32, yhi = y, None
    out = _empty_planes(dy.shape, dy.device)
    _launch("lfs2_relu_bwd_planes", _p(dy), _p(y32), _p(yhi), _p(out.hi), _p(out.lo), _p(db), rows, cols, float(scale),
            _s(), tag="lfs2_relu_bwd", nbytes=(10.0 if yhi is not None else 12.0) * dy.numel())
    return out


class WeightPrepPlan:
    """Persistent buffers + device table of one lfs2_weight_planes_batched launch: sources = [(fp32 (rows, cols) tensor,
    want planes of W, want planes of W^T)].  `planes[i]` / `planes_t[i]` are the Planes the launch fills (or None)."""

    def __init__(self, sources):
        import numpy as np

        dev = sources[0][0].device
        pad = lambda n: (n + 63) // 64 * 64   # every plane starts on a 128-byte boundary (TMA operands)
        total = sum(pad(w.numel()) * (2 * int(p) + 2 * int(t)) for w, p, t in sources)
        self.buf = torch.empty(total, device=dev, dtype=torch.bfloat16)
        dt = np.dtype([("src", "<u8"), ("hi", "<u8"), ("lo", "<u8"), ("hi_t", "<u8"), ("lo_t", "<u8"), ("rows", "<i4"),
                       ("cols", "<i4"), ("tile_begin", "<i4"), ("reserved", "<i4")])
        tab = np.zeros(len(sources), dtype=dt)
        self.planes, self.planes_t, self.sources = [], [], [w for w, _, _ in sources]
        off = tiles = 0
        for i, (w, want_p, want_t) in enumerate(sources):
            _chk(w, torch.float32, "weight_prep source", 2)
            rows, cols = w.shape
            n = rows * cols

            def take(shape):
                nonlocal off
                v = self.buf[off:off + n].view(shape)
                off += pad(n)
                return v

            pl = Planes(take((rows, cols)), take((rows, cols))) if want_p else None
            pt = Planes(take((cols, rows)), take((cols, rows))) if want_t else None
            self.planes.append(pl)
            self.planes_t.append(pt)
            tab[i] = (w.data_ptr(), pl.hi.data_ptr() if pl else 0, pl.lo.data_ptr() if pl else 0,
                      pt.hi.data_ptr() if pt else 0, pt.lo.data_ptr() if pt else 0, rows, cols, tiles, 0)
            tiles += ((rows + 31) // 32) * ((cols + 31) // 32)
        self.n, self.tiles = len(sources), tiles
        self.table = torch.from_numpy(tab.view(np.uint8).copy()).to(dev)
        self.signature = tuple((w.data_ptr(), tuple(w.shape), p, t) for w, p, t in sources)

    def run(self):
        """(re)fill every output from the sources' current values: one launch"""
        _launch("lfs2_weight_planes_batched", _p(self.table), self.n, self.tiles, _s(), tag="lfs2_weight_prep",
                nbytes=float(sum(w.numel() for w in self.sources)) * 4.0 + 2.0 * self.buf.numel())


def add_(dst, src):
    _launch("lfs2_add_inplace", _p(dst), _p(src), dst.numel(), _s(), nbytes=12.0 * dst.numel())
    drop_planes(dst)
    return dst


def transpose(w):
    """(rows, cols) fp32 -> (cols, rows)"""
    _chk(w, torch.float32, "transpose input", 2)
    out = torch.empty(w.shape[1], w.shape[0], device=w.device, dtype=torch.float32)
    _launch("lfs2_transpose", _p(w), _p(out), w.shape[0], w.shape[1], _s())
    return out


def dwconv1d_bwd_w_(dwt, dbias, dy, x):
    b, t, d = x.shape
    _launch("lfs2_dwconv1d_bwd_w", _p(dy), _p(x), _p(dwt), _p(dbias), b, t, d, dwt.shape[0], _s(),
            flops=2.0 * b * t * d * dwt.shape[0], nbytes=8.0 * b * t * d)


def attention_lse(qkv, kpm, nhead):
    """-> (ctx (B,T,d), lse (B,nhead,T))"""
    _chk(qkv, torch.float32, "qkv", 3)
    b, t, d3 = qkv.shape
    d = d3 // 3
    ctx = torch.empty(b, t, d, device=qkv.device, dtype=torch.float32)
    lse = torch.empty(b, nhead, t, device=qkv.device, dtype=torch.float32)
    _launch("lfs2_attention_lse", _p(qkv), _p(kpm), _p(ctx), _p(lse), b, t, d, nhead, _s(), tag="lfs2_attention",
            flops=4.0 * d * t * t * b, nbytes=4.0 * (qkv.numel() + ctx.numel()))
    return ctx, lse


def attention_bwd(qkv, ctx, dctx, lse, kpm, nhead):
    b, t, d3 = qkv.shape
    d = d3 // 3
    dqkv = torch.empty_like(qkv)
    ws = torch.empty(max(4, _lib.lib().lfs2_attention_bwd_workspace_bytes(b, t, nhead)), device=qkv.device,
                     dtype=torch.uint8)
    _launch("lfs2_attention_bwd", _p(qkv), _p(ctx), _p(dctx), _p(lse), _p(kpm), _p(dqkv), _p(ws), b, t, d, nhead, _s(),
            flops=14.0 * d * t * t * b, nbytes=4.0 * (2 * qkv.numel() + 2 * ctx.numel()))
    return dqkv


def length_regulate_train(x, durations, max_length, frames=None):
    """length_regulate that also returns the prefix sums its backward needs; frames = (l, cap) as in length_regulate"""
    scan = length_regulate_scan(durations, x.shape[:2])
    out, mask = length_regulate(x, durations, max_length, scan=scan, frames=frames)
    return out, mask, scan[0]


def length_regulate_bwd(dout, cum):
    b, l, d = dout.shape
    tp = cum.shape[1]
    dx = torch.empty(b, tp, d, device=dout.device, dtype=torch.float32)
    _launch("lfs2_length_regulate_bwd", _p(dout), _p(cum), _p(dx), b, tp, l, d, _s(),
            nbytes=4.0 * d * (b * l + b * tp))
    return dx


def embedding_bwd_(demb, dx, idx, skip_idx=-1):
    _chk(idx, torch.int64, "embedding indices")
    d = dx.shape[-1]
    _launch("lfs2_embedding_bwd", _p(dx), _p(idx), _p(demb), dx.numel() // d, d, demb.shape[0], skip_idx, _s(),
            nbytes=4.0 * dx.numel())


def bucket_embed_add_oop(x_in, val, std, mean, bins, emb, idx_forced=None):
    """-> (x_in + emb[idx], idx)"""
    _chk(x_in, torch.float32, "x", 3)
    b, t, d = x_in.shape
    out = torch.empty_like(x_in)
    idx_out = torch.empty(b, t, device=x_in.device, dtype=torch.int64)
    _launch("lfs2_bucket_embed_add_oop", _p(x_in), _p(out), _p(val), float(std), float(mean), _p(bins), emb.shape[0],
            _p(emb), _p(idx_forced), _p(idx_out), _p(None), 0, b * t, d, _s(), tag="lfs2_bucket_embed_add",
            nbytes=8.0 * b * t * d)
    return out, idx_out


def rowdot_mask_bwd(dout, z, w, mask, dw, db):
    f = z.shape[-1]
    dz = torch.empty_like(z)
    _launch("lfs2_rowdot_mask_bwd", _p(dout), _p(z), _p(w), _p(mask), _p(dz), _p(dw), _p(db), z.numel() // f, f, _s(),
            nbytes=8.0 * z.numel())
    return dz


def sum_over_time_(out, dx):
    b, t, d = dx.shape
    _launch("lfs2_sum_over_time", _p(dx), _p(out), b, t, d, _s(), nbytes=4.0 * dx.numel())


def fold_pw(w21, w20, b20, b21):
    """-> (w_eff (d_out, F), b_eff (d_out)) of conv2.1 . conv2.0 (grouped 1x1)"""
    d_out, f = w21.shape[0], w21.shape[1]
    g = w20.shape[1]
    w_eff = torch.empty(d_out, f, device=w21.device, dtype=torch.float32)
    b_eff = torch.empty(d_out, device=w21.device, dtype=torch.float32)
    _launch("lfs2_fold_pw_fwd", _p(w21), _p(w20), _p(b20), _p(b21), _p(w_eff), _p(b_eff), d_out, f // g, g, _s())
    return w_eff, b_eff


def fold_pw_bwd_(dw_eff, db_eff, w21, w20, b20, dw21, dw20, db20, db21):
    d_out, f = w21.shape[0], w21.shape[1]
    g = w20.shape[1]
    _launch("lfs2_fold_pw_bwd", _p(dw_eff), _p(db_eff), _p(w21), _p(w20), _p(b20), _p(dw21), _p(dw20), _p(db20),
            _p(db21), d_out, f // g, g, _s())


def masked_loss(pred, target, pad_mask, kind, weight, loss_out, total_out, want_grad=True, target_i64=None):
    """loss_out / total_out: 1-element fp32 device tensors (views).  -> d(weight*loss)/d pred or None"""
    _chk(pred, torch.float32, "loss prediction")
    rows = pad_mask.numel()
    inner = pred.numel() // rows
    if target_i64 is not None:
        _chk(target_i64, torch.int64, "integer loss target")
        if target_i64.numel() != pred.numel():
            raise ValueError("masked_loss: target shape mismatch")
    else:
        _chk(target, torch.float32, "loss target")
        if target.numel() != pred.numel():
            raise ValueError(f"masked_loss: target {tuple(target.shape)} vs prediction {tuple(pred.shape)}")
    dpred = torch.empty_like(pred) if want_grad else None
    ws = torch.empty(2, device=pred.device, dtype=torch.float32)
    _launch("lfs2_masked_loss", _p(pred), _p(target), _p(target_i64), _p(pad_mask), rows, inner,
            {"l1": 0, "mse": 1}[kind], float(weight), _p(loss_out), _p(total_out), _p(dpred), _p(ws), _s(),
            nbytes=4.0 * pred.numel() * (3 if want_grad else 2))
    return dpred


def sumsq_(out, x):
    _launch("lfs2_sumsq", _p(x), _p(out), x.numel(), _s(), nbytes=4.0 * x.numel())


def adamw_step_(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0, max_norm=0.0, gnorm_sq=None,
                zero_grad=True):
    for t_ in (p, g, m, v):
        _chk(t_, torch.float32, "adamw flat buffer", 1)
    _launch("lfs2_adamw_step", _p(p), _p(g), _p(m), _p(v), p.numel(), float(lr), float(beta1), float(beta2), float(eps),
            float(weight_decay), int(step), float(grad_scale), float(max_norm), _p(gnorm_sq), int(zero_grad), _s(),
            nbytes=28.0 * p.numel())


def scale_by_(x, scalar):
    """x *= scalar (a 0-dim / 1-element fp32 device tensor), no host sync"""
    _launch("lfs2_scale_by", _p(x), _p(scalar), x.numel(), _s(), nbytes=8.0 * x.numel())
    return x


def _operand(t, mn_major, col0=0, hstride=0, per_z=False):
    """lfs2_operand for a contiguous bf16 tensor viewed as (d2, d1, d0)"""
    if t.dim() == 2:
        d2, d1, d0 = 1, t.shape[0], t.shape[1]
    else:
        d2, d1, d0 = t.shape[0], t.shape[1], t.shape[2]
    return _lib.Operand(int(mn_major), d0, d1, d2, col0, hstride, int(per_z))


def gemm_tc2(a, a_op, b, b_op, c, ldc, m, n, k, nbatch=1, nhead=1, c_bstride=0, c_hstride=0, c_offset=0, npass=3,
             accumulate=False, tag=None):
    """C[z] (+)= A[z] . B[z]^T on tcgen05 with per-operand majorness (see lfs2_gemm_tc2)."""
    for x_ in (a.hi, b.hi):
        _chk(x_, torch.bfloat16, "gemm_tc2 operand plane")
    _launch("lfs2_gemm_tc2", _p(a.hi), _p(a.lo if npass == 3 else None), ctypes.byref(a_op), _p(b.hi),
            _p(b.lo if npass == 3 else None), ctypes.byref(b_op), ctypes.c_void_p(c.data_ptr() + 4 * c_offset), ldc,
            c_bstride, c_hstride, m, n, k, nbatch, nhead, npass, int(accumulate), _s(), tag=tag or "gemm_tc2",
            flops=2.0 * m * n * k * nbatch * nhead, passes=npass, nbytes=4.0 * nbatch * nhead * (m * k + n * k + m * n))


def wgrad_tc_ok(n, k):
    """shapes the tcgen05 weight-gradient kernel covers (else the CUDA-core gemm_tn runs)"""
    return n % 8 == 0 and k % 8 == 0 and n >= 32 and k >= 32


def gemm_wgrad_tc_(dw, dy, x, npass=3, tag=None):
    """dw (n, k) += dy (m, n)^T . x (m, k) on the tensor cores (both operands MN-major: the contraction runs
    over the rows of dy and x, no transposed copies)."""
    n, k = dw.shape
    m = dy.hi.numel() // n
    dy2 = Planes(dy.hi.view(m, n), dy.lo.view(m, n) if dy.lo is not None else None)
    x2 = Planes(x.hi.view(m, k), x.lo.view(m, k) if x.lo is not None else None)
    gemm_tc2(dy2, _operand(dy2.hi, True), x2, _operand(x2.hi, True), dw, k, n, k, m, npass=npass, accumulate=True,
             tag=tag or f"wgrad_tc_n{n}_k{k}")


def dropout_(x, p, seed, site, out=None):
    """x * keep / (1 - p) with keep = Philox(seed, site, element index); in place unless `out` is given.
    Calling it on a gradient with the same (seed, site) is the backward pass."""
    _chk(x, torch.float32, "dropout input")
    y = x if out is None else out
    _launch("lfs2_dropout", _p(x), _p(y), x.numel(), float(p), int(seed), int(site), _s(), nbytes=8.0 * x.numel())
    drop_planes(y)
    return y


def dropout_planes(pl, p, seed, site):
    out = Planes(torch.empty_like(pl.hi), torch.empty_like(pl.lo) if pl.lo is not None else None)
    _launch("lfs2_dropout_planes", _p(pl.hi), _p(pl.lo), _p(out.hi), _p(out.lo), pl.hi.numel(), float(p), int(seed),
            int(site), _s(), tag="lfs2_dropout", nbytes=(8.0 if pl.lo is not None else 4.0) * pl.hi.numel())
    return out


def attention_mat_fwd(qkv, kpm, nhead, npass=3, drop=None):
    """GEMM-decomposed attention forward on qkv Planes (B,T,3d): -> (ctx fp32 (B,T,d), P Planes (Z,T,Tp), lse (Z,T));
    drop = (p, seed, site): attention-probability dropout, then the 2nd result is (P, P o mask/(1-p))."""
    b, t, d3 = qkv.shape
    d = d3 // 3
    dh = d // nhead
    z = b * nhead
    tp = (t + 7) // 8 * 8
    dev = qkv.hi.device
    s = torch.empty(z, t, tp, device=dev, dtype=torch.float32)
    q_op = _operand(qkv.hi, False, col0=0, hstride=dh)
    k_op = _operand(qkv.hi, False, col0=d, hstride=dh)
    gemm_tc2(qkv, q_op, qkv, k_op, s, tp, t, t, dh, nbatch=b, nhead=nhead, c_bstride=nhead * t * tp, c_hstride=t * tp,
             npass=npass, tag="attn_qk_gemm")
    p = Planes(torch.empty(z, t, tp, device=dev, dtype=torch.bfloat16),
               torch.empty(z, t, tp, device=dev, dtype=torch.bfloat16) if npass == 3 else None)
    lse = torch.empty(z, t, device=dev, dtype=torch.float32)
    pm = p
    if drop is not None:
        pm = Planes(torch.empty_like(p.hi), torch.empty_like(p.lo) if p.lo is not None else None)
    dp_, dseed, dsite = drop if drop is not None else (0.0, 0, 0)
    _launch("lfs2_attn_softmax_planes_drop", _p(s), _p(kpm), _p(p.hi), _p(p.lo), _p(pm.hi if drop is not None else None),
            _p(pm.lo if drop is not None else None), _p(lse), b, nhead, t, tp, float(dh) ** -0.5, float(dp_), int(dseed),
            int(dsite), _s(), tag="lfs2_attn_softmax_planes",
            nbytes=((8.0 if npass == 3 else 6.0) + (0.0 if drop is None else (4.0 if npass == 3 else 2.0))) * z * t * tp)
    del s
    ctx = torch.empty(b, t, d, device=dev, dtype=torch.float32)
    p_op = _operand(pm.hi, False, per_z=True)
    v_op = _operand(qkv.hi, True, col0=2 * d, hstride=dh)
    gemm_tc2(pm, p_op, qkv, v_op, ctx, d, t, dh, t, nbatch=b, nhead=nhead, c_bstride=t * d, c_hstride=dh, npass=npass,
             tag="attn_pv_gemm")
    return ctx, (p if drop is None else (p, pm)), lse


def attention_mat_bwd(qkv, p, ctx, dctx, nhead, npass=3, drop=None):
    """backward of attention_mat_fwd: -> dqkv fp32 (B,T,3d) = [dq | dk | dv]"""
    pm = p
    if drop is not None:
        p, pm = p
    b, t, d3 = qkv.shape
    d = d3 // 3
    dh = d // nhead
    z = b * nhead
    tp = p.hi.shape[-1]
    dev = qkv.hi.device
    do = split_bf16(dctx)
    delta = torch.empty(z, t, device=dev, dtype=torch.float32)
    _launch("lfs2_attn_delta", _p(dctx), _p(ctx), _p(delta), b, t, d, nhead, _s(), nbytes=8.0 * dctx.numel())
    dp = torch.empty(z, t, tp, device=dev, dtype=torch.float32)
    do_k = _operand(do.hi, False, col0=0, hstride=dh)
    v_k = _operand(qkv.hi, False, col0=2 * d, hstride=dh)
    gemm_tc2(do, do_k, qkv, v_k, dp, tp, t, t, dh, nbatch=b, nhead=nhead, c_bstride=nhead * t * tp, c_hstride=t * tp,
             npass=npass, tag="attn_dp_gemm")
    ds = Planes(torch.empty(z, t, tp, device=dev, dtype=torch.bfloat16),
                torch.empty(z, t, tp, device=dev, dtype=torch.bfloat16) if npass == 3 else None)
    # O = (P o M).V  =>  dP_eff = M o (dO.V^T), applied inside the kernel; delta = rowsum(dO o O) still holds
    dp_, dseed, dsite = drop if drop is not None else (0.0, 0, 0)
    _launch("lfs2_attn_ds_planes_drop", _p(p.hi), _p(p.lo), _p(dp), _p(delta), _p(ds.hi), _p(ds.lo), b, nhead, t, tp,
            float(dh) ** -0.5, float(dp_), int(dseed), int(dsite), _s(), tag="lfs2_attn_ds_planes",
            nbytes=(12.0 if npass == 3 else 8.0) * z * t * tp)
    del dp
    dqkv = torch.empty(b, t, d3, device=dev, dtype=torch.float32)
    common = dict(nbatch=b, nhead=nhead, c_bstride=t * d3, c_hstride=dh, npass=npass)
    # dV = P^T . dO ; dK = dS^T . (scale folded into dS) Q ; dQ = dS . K
    gemm_tc2(pm, _operand(pm.hi, True, per_z=True), do, _operand(do.hi, True, col0=0, hstride=dh), dqkv, d3, t, dh, t,
             c_offset=2 * d, tag="attn_dv_gemm", **common)
    gemm_tc2(ds, _operand(ds.hi, True, per_z=True), qkv, _operand(qkv.hi, True, col0=0, hstride=dh), dqkv, d3, t, dh, t,
             c_offset=d, tag="attn_dk_gemm", **common)
    gemm_tc2(ds, _operand(ds.hi, False, per_z=True), qkv, _operand(qkv.hi, True, col0=d, hstride=dh), dqkv, d3, t, dh,
             t, c_offset=0, tag="attn_dq_gemm", **common)
    return dqkv
