"""Host-side mirror of ``litfass.third_party.hifigan`` (reference litfass/third_party/hifigan/models.py:20-174 and
__init__.py:18-42): the HiFi-GAN v1 generator that turns the mel frames of the path into a waveform -- the step right
behind ``FastSpeech2.forward`` in every synthesis call (synthesis/generator.py:170, fastspeech2.py:917-918).

Same class names, constructor signatures and state_dict keys as the reference (``weight_norm``'s ``weight_g`` /
``weight_v`` pairs included, so ``generator_universal.pth.tar`` loads unchanged); the torch modules are parameter
containers only.  ``Generator.forward`` runs on the kernels of liblfs2.so:

* activations travel channels-last as bf16 hi/lo planes (x = hi + lo);
* every Conv1d -- ``conv_pre`` (80 -> 512, k 7), the 36 dilated / plain ResBlock convolutions (k 3 / 7 / 11,
  dilation 1 / 3 / 5) -- is ``lfs2_gemm_tc_ex``: k tap-shifted tcgen05 GEMMs into one TMEM accumulator, TMA zero fill
  = the "same" padding, bias + leaky-ReLU or bias + residual fused in the epilogue;
* every ConvTranspose1d(C_in -> C_out, kernel 2u, stride u, padding u/2) is the SAME kernel: in polyphase form output
  sample t*u + r only sees inputs t-1, t, t+1, so it is a 3-tap convolution onto u*C_out columns whose row-major
  (T, u*C_out) result IS the (T*u, C_out) upsampled tensor -- no zero-stuffing, no scatter;
* the element-wise stages between them (leaky ReLU in front of a ResBlock conv, the average of the three ResBlocks,
  leaky ReLU + conv_post (32 -> 1) + tanh) are streaming kernels (csrc/vocoder.cu).

Ragged batches: ``forward(mel, lengths)`` processes utterances of different lengths in one padded batch; every kernel
writes zeros on the rows past an utterance's end, so each utterance sees exactly the zero padding a stand-alone call
would give it and its samples equal the per-utterance result.
"""
import json
import os

import torch
from torch import nn
from torch.nn import Conv1d, ConvTranspose1d
from torch.nn.utils import remove_weight_norm, weight_norm

from . import ops

LRELU_SLOPE = 0.1


class AttrDict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


# reference third_party/hifigan/config.json (the fields the generator reads)
DEFAULT_CONFIG = AttrDict(
    resblock="1", upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
    resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], num_mels=80,
    sampling_rate=22050, hop_size=256)


def init_weights(m, mean=0.0, std=0.01):
    if m.__class__.__name__.find("Conv") != -1:
        m.weight.data.normal_(mean, std)


def get_padding(kernel_size, dilation=1):
    return int((kernel_size * dilation - dilation) / 2)


def _effective_weight(conv):
    """the convolution's weight: g * v / |v| while weight_norm is attached (torch's dim=0 convention), else .weight"""
    if hasattr(conv, "weight_g"):
        return torch._weight_norm(conv.weight_v, conv.weight_g, 0)
    return conv.weight


def _conv_planes(conv, pad_in_to=None):
    """Conv1d weight (C_out, C_in, k) -> tap-major (C_out, k * C_in') bf16 hi/lo planes, C_in zero-padded to C_in'"""
    w = _effective_weight(conv).detach().float()
    n, c, k = w.shape
    cp = pad_in_to or c
    wp = torch.zeros(n, k, cp, device=w.device, dtype=torch.float32)
    wp[:, :, :c] = w.permute(0, 2, 1)
    return ops.split_bf16(wp.reshape(n, k * cp).contiguous())


def _upsample_planes(conv):
    """ConvTranspose1d weight (C_in, C_out, K), stride u, padding p, K = 2u, p = u/2 -> the equivalent 3-tap Conv1d
    onto u * C_out columns (column r * C_out + co = output phase r, channel co):
        out[t*u + r, co] = sum_{j in {1,0,-1}} x[t - j] . W[:, co, j*u + r + p]   where 0 <= j*u + r + p < K
    as tap-major planes (u * C_out, 3 * C_in); tap index 0 / 1 / 2 = input row t-1 / t / t+1."""
    w = _effective_weight(conv).detach().float()
    c_in, c_out, kk = w.shape
    u, p = conv.stride[0], conv.padding[0]
    if kk != 2 * u or 2 * p != kk - u or conv.output_padding[0] != 0 or conv.dilation[0] != 1:
        raise NotImplementedError("ConvTranspose1d: only kernel = 2 * stride, padding = (kernel - stride) / 2 "
                                  "(every upsampler of the reference's config)")
    weq = torch.zeros(u, c_out, 3, c_in, device=w.device, dtype=torch.float32)
    for tap, j in enumerate((1, 0, -1)):
        for r in range(u):
            k = j * u + r + p
            if 0 <= k < kk:
                weq[r, :, tap, :] = w[:, :, k].t()
    bias = conv.bias.detach().float().repeat(u).contiguous()
    return ops.split_bf16(weq.reshape(u * c_out, 3 * c_in).contiguous()), bias


class ResBlock(nn.Module):
    """reference models.py:20-110 (ResBlock type "1")"""

    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3, 5)):
        super().__init__()
        self.h = h
        self.kernel_size = kernel_size
        self.dilation = tuple(dilation)
        self.convs1 = nn.ModuleList([
            weight_norm(Conv1d(channels, channels, kernel_size, 1, dilation=d, padding=get_padding(kernel_size, d)))
            for d in dilation])
        self.convs1.apply(init_weights)
        self.convs2 = nn.ModuleList([
            weight_norm(Conv1d(channels, channels, kernel_size, 1, dilation=1, padding=get_padding(kernel_size, 1)))
            for _ in dilation])
        self.convs2.apply(init_weights)

    def remove_weight_norm(self):
        for conv in list(self.convs1) + list(self.convs2):
            remove_weight_norm(conv)

    def forward_planes(self, x, xa, pack, row_mask, npass, row_limit=None):
        """x: Planes (B, T, C); xa = leaky_relu(x) if the caller already has it; -> Planes"""
        k = self.kernel_size
        for m, d in enumerate(self.dilation):
            a = xa if (m == 0 and xa is not None) else ops.lrelu_planes(x, LRELU_SLOPE)
            w1, w2 = pack[m]
            hh = ops.gemm_tc(a, w1, self.convs1[m].bias, taps=k, dilation=d, leaky_slope=LRELU_SLOPE, out="planes",
                             npass=npass, row_mask=row_mask, row_limit=row_limit, tag="hifigan_resblock_conv1")
            # (x + conv2(.): the residual rides the tensor core as hi/lo identity slabs, which needs the 3-pass stage
            #  layout -- the residual stream keeps fp32 precision in "bf16" mode as well)
            x = ops.gemm_tc(hh, w2, self.convs2[m].bias, taps=k, residual=x, out="planes", npass=3,
                            row_mask=row_mask, row_limit=row_limit, tag="hifigan_resblock_conv2")
        return x


class Generator(nn.Module):
    """reference models.py:112-174.  forward(x (B, num_mels, T) fp32[, lengths (B)]) -> (B, 1, T * prod(upsample_rates))"""

    def __init__(self, h):
        super().__init__()
        self.h = h
        if str(getattr(h, "resblock", "1")) != "1":
            raise NotImplementedError("ResBlock type 2 (the reference implements type 1 only)")
        self.num_kernels = len(h.resblock_kernel_sizes)
        self.num_upsamples = len(h.upsample_rates)
        if self.num_kernels != 3:
            raise NotImplementedError("the multi-receptive-field average is implemented for 3 ResBlocks per stage")
        self.conv_pre = weight_norm(Conv1d(getattr(h, "num_mels", 80), h.upsample_initial_channel, 7, 1, padding=3))
        self.ups = nn.ModuleList()
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            self.ups.append(weight_norm(ConvTranspose1d(h.upsample_initial_channel // (2 ** i),
                                                        h.upsample_initial_channel // (2 ** (i + 1)), k, u,
                                                        padding=(k - u) // 2)))
        self.resblocks = nn.ModuleList()
        ch = h.upsample_initial_channel
        for i in range(len(self.ups)):
            ch = h.upsample_initial_channel // (2 ** (i + 1))
            for k, d in zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes):
                self.resblocks.append(ResBlock(h, ch, k, d))
        self.conv_post = weight_norm(Conv1d(ch, 1, 7, 1, padding=3))
        self.ups.apply(init_weights)
        self.conv_post.apply(init_weights)
        self._pack_key = None
        self._pack = None

    compute_mode = "fp32"   # "fp32": split-bf16 operands, 3 tensor-core passes per product; "bf16": one pass

    def remove_weight_norm(self):
        print("Removing weight norm...")
        for conv in self.ups:
            remove_weight_norm(conv)
        for block in self.resblocks:
            block.remove_weight_norm()
        remove_weight_norm(self.conv_pre)
        remove_weight_norm(self.conv_post)

    # -- kernel-side weight layouts, rebuilt when any parameter changes ---------------------------------------
    def _packed(self):
        params = list(self.parameters())
        key = (ops.WEIGHTS_EPOCH,) + tuple((p.data_ptr(), p._version) for p in params)
        if key != self._pack_key:
            with torch.no_grad():
                c_in = self.conv_pre.in_channels
                pk = {"c_in_padded": (c_in + 31) // 32 * 32}
                pk["pre"] = _conv_planes(self.conv_pre, pad_in_to=pk["c_in_padded"])
                pk["ups"] = [_upsample_planes(conv) for conv in self.ups]
                pk["res"] = [[(_conv_planes(c1), _conv_planes(c2)) for c1, c2 in zip(b.convs1, b.convs2)]
                             for b in self.resblocks]
                wpost = _effective_weight(self.conv_post).detach().float()   # (1, C, k)
                pk["post"] = wpost[0].t().contiguous().reshape(-1)           # tap-major (k * C)
            self._pack, self._pack_key = pk, key
        return self._pack

    def forward(self, x, lengths=None):
        """x: (B, num_mels, T) fp32 on a CUDA device (channels-first, as the reference takes it); lengths: optional
        (B) valid frame counts of a zero-padded ragged batch -> (B, 1, T * hop) fp32, zero past lengths * hop."""
        if not (torch.is_tensor(x) and x.is_cuda):
            raise ops._lib.Lfs2Error("hifigan.Generator.forward needs a CUDA tensor: there is no CPU path")
        if x.dim() != 3 or x.shape[1] != self.conv_pre.in_channels:
            raise ValueError(f"expected (B, {self.conv_pre.in_channels}, T), got {tuple(x.shape)}")
        npass = 3 if self.compute_mode == "fp32" else 1
        pk = self._packed()
        x = x.contiguous().float()
        bsz, _, t = x.shape
        dev = x.device
        len32 = None
        if lengths is not None:
            len32 = torch.as_tensor(lengths, device=dev).to(torch.int32).contiguous()

        def row_mask(scale):  # (B, t * scale) bool, True = past the utterance's end at this stage's resolution
            if len32 is None:
                return None
            return (torch.arange(t * scale, device=dev)[None, :] >= (len32 * scale)[:, None]).contiguous()

        def row_limit(scale):
            """128-row tiles that start at or after length + 32 rows are skipped altogether: the rows they would write
            are zeros nobody needs -- a valid sample reads at most 25 rows (k 11, dilation 5) past its utterance's end,
            rows the last kept tile has written as zeros -- and the element-wise stages may carry anything there"""
            if len32 is None:
                return None
            return ((len32 * scale).contiguous(), 32, {})

        xp = ops.mel_to_planes(x, len32, pk["c_in_padded"])
        mask, lim = row_mask(1), row_limit(1)
        # conv_pre, with the first stage's leaky ReLU in its epilogue (nothing else reads the un-activated tensor)
        hcur = ops.gemm_tc(xp, pk["pre"], self.conv_pre.bias, taps=7, leaky_slope=LRELU_SLOPE, out="planes", npass=npass,
                           row_mask=mask, row_limit=lim, tag="hifigan_conv_pre")
        scale = 1
        for i, up in enumerate(self.ups):
            u = up.stride[0]
            w_up, b_up = pk["ups"][i]
            y = ops.gemm_tc(hcur, w_up, b_up, taps=3, out="planes", npass=npass, row_mask=mask, row_limit=lim,
                            tag="hifigan_upsample")
            scale *= u
            c_out = up.out_channels
            xu = ops.Planes(y.hi.view(bsz, t * scale, c_out), y.lo.view(bsz, t * scale, c_out))
            mask, lim = row_mask(scale), row_limit(scale)
            xa = ops.lrelu_planes(xu, LRELU_SLOPE)  # shared by the first convolution of the stage's three ResBlocks
            outs = [self.resblocks[i * self.num_kernels + j].forward_planes(xu, xa, pk["res"][i * self.num_kernels + j],
                                                                            mask, npass, lim)
                    for j in range(self.num_kernels)]
            del xa, xu, y
            last = i + 1 == self.num_upsamples
            # x = (r1 + r2 + r3) / 3, then the next stage's leaky_relu(0.1) -- or F.leaky_relu's default 0.01 in front of
            # conv_post (models.py:167)
            hcur = ops.mean3_lrelu_planes(outs[0], outs[1], outs[2], 0.01 if last else LRELU_SLOPE)
            del outs
        lens_out = None if len32 is None else (len32 * scale).contiguous()
        wav = ops.conv_post_tanh(hcur, pk["post"], self.conv_post.bias, lens_out, 1.0)
        return wav.unsqueeze(1)


def find_checkpoint(model="universal"):
    """generator_<model>.pth.tar of an installed reference (litfass/third_party/hifigan/), or None: the 56 MB weight
    files are the reference's data and are not redistributed here"""
    try:
        import importlib.util

        spec = importlib.util.find_spec("litfass")
        roots = list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []
    except Exception:  # noqa: BLE001
        roots = []
    for root in roots:
        p = os.path.join(root, "third_party", "hifigan", f"generator_{model}.pth.tar")
        if os.path.exists(p):
            return p
    return None


class Synthesiser:
    """reference third_party/hifigan/__init__.py:18-42: mel (T, num_mels) -> int16 samples (1, T * hop).
    ``checkpoint`` (path) and ``config`` (dict or path) default to the reference's bundled files when ``litfass`` is
    importable; ``batch()`` vocodes a list of mels of different lengths in one padded launch sequence."""

    def __init__(self, device="cuda:0", model="universal", checkpoint=None, config=None):
        if isinstance(config, (str, os.PathLike)):
            with open(config) as f:
                config = json.load(f)
        h = AttrDict(config) if config is not None else AttrDict(DEFAULT_CONFIG)
        vocoder = Generator(h)
        path = checkpoint or find_checkpoint(model)
        if path is None:
            raise FileNotFoundError(
                f"generator_{model}.pth.tar not found: pass checkpoint=<path> (the weights ship with the reference under "
                "litfass/third_party/hifigan/)")
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        vocoder.load_state_dict(ckpt["generator"])
        vocoder.eval()
        vocoder.remove_weight_norm()
        self.device = device
        vocoder.to(self.device)
        self.vocoder = vocoder

    def __call__(self, mel):
        mel = torch.unsqueeze(mel.T, 0)
        with torch.no_grad():
            wav = self.vocoder(mel.to(self.device).float())
        return (wav.squeeze(1).cpu().detach().numpy() * 32768.0).astype("int16")

    def batch(self, mels):
        """mels: list of (T_i, num_mels) tensors -> list of int16 arrays (T_i * hop,)"""
        lens = [int(m.shape[0]) for m in mels]
        t = max(lens)
        x = torch.zeros(len(mels), mels[0].shape[1], t, device=self.device)
        for i, m in enumerate(mels):
            x[i, :, : lens[i]] = m.to(self.device).float().T
        with torch.no_grad():
            wav = self.vocoder(x, torch.tensor(lens))
        hop = wav.shape[-1] // t
        wav = (wav.squeeze(1).cpu().numpy() * 32768.0).astype("int16")
        return [wav[i, : lens[i] * hop] for i in range(len(mels))]
