"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's HiFi-GAN generator forward
(litfass/third_party/hifigan/models.py:20-174, Synthesiser.__call__ at __init__.py:36-42), written against a plain
{name: tensor} state dict with weight_norm already folded (``<conv>.weight`` / ``<conv>.bias``; ``fold_weight_norm``
does that for a checkpoint that still carries ``weight_g`` / ``weight_v``).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this file.  PINNED against the
unmodified reference ``Generator`` run in the authoring container: seeded weights (tests/golden/hifigan_small.pt,
written by oracle/make_goldens_hifigan.py) and -- when /root/reference is present -- its bundled
generator_universal.pth.tar (tests/test_oracle_hifigan.py).
"""
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # models.py:7

# third_party/hifigan/config.json
CONFIG = dict(upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=512,
              resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], num_mels=80)


def fold_weight_norm(sd):
    """{..weight_g, ..weight_v} -> {..weight}: w = g * v / |v| with the norm over every dim but 0 (torch weight_norm
    default dim=0, which the reference uses for Conv1d AND ConvTranspose1d, models.py:28-60,119-133)"""
    out = {}
    for k, v in sd.items():
        if k.endswith("weight_g"):
            base = k[: -len("weight_g")]
            vv = sd[base + "weight_v"]
            norm = vv.reshape(vv.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (vv.dim() - 1)))
            out[base + "weight"] = v * vv / norm
        elif not k.endswith("weight_v"):
            out[k] = v
    return out


def resblock(x, sd, pre, kernel_size, dilations):
    """ResBlock.forward, models.py:84-91"""
    for m, d in enumerate(dilations):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, sd[f"{pre}convs1.{m}.weight"], sd[f"{pre}convs1.{m}.bias"], dilation=d,
                      padding=(kernel_size * d - d) // 2)
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, sd[f"{pre}convs2.{m}.weight"], sd[f"{pre}convs2.{m}.bias"], padding=(kernel_size - 1) // 2)
        x = xt + x
    return x


def generator(sd, mel, cfg=None):
    """Generator.forward, models.py:150-171: mel (B, num_mels, T) -> (B, 1, T * prod(upsample_rates))"""
    cfg = cfg or CONFIG
    nk = len(cfg["resblock_kernel_sizes"])
    x = F.conv1d(mel, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, sd[f"ups.{i}.weight"], sd[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j, (ks, ds) in enumerate(zip(cfg["resblock_kernel_sizes"], cfg["resblock_dilation_sizes"])):
            r = resblock(x, sd, f"resblocks.{i * nk + j}.", ks, ds)
            xs = r if xs is None else xs + r
        x = xs / nk
    x = F.leaky_relu(x)  # default slope 0.01 (models.py:167)
    x = F.conv1d(x, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)
    return torch.tanh(x)


def synthesise(sd, mel, cfg=None):
    """Synthesiser.__call__ (__init__.py:36-42): mel (T, num_mels) -> int16 numpy (1, T * hop)"""
    wav = generator(sd, mel.T.unsqueeze(0).float(), cfg)
    return (wav.squeeze(1).detach().numpy() * 32768.0).astype("int16")
