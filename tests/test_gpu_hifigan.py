"""HiFi-GAN generator on the GPU (lightningfastspeech2_b200.hifigan) against the reference goldens and
the oracle.  Tolerance: 1e-3 abs on the float waveform in [-1, 1] in "fp32" mode (split-bf16 x3 tensor-core passes; the
same budget north_star gives the mel), 3e-2 in "bf16" mode; masks / lengths / shapes exact."""
import os

import numpy as np
import pytest
import torch

from lightningfastspeech2_b200 import hifigan, synthetic
from oracle import hifigan_oracle as HO

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {"fp32": 1e-3, "bf16": 3e-2}


def build(cfg, seed, mode="fp32"):
    gen = hifigan.Generator(hifigan.AttrDict(cfg))
    gen.remove_weight_norm()
    sd = synthetic.hifigan_state_dict(cfg, seed=seed)
    gen.load_state_dict(sd, strict=True)
    gen.compute_mode = mode
    return gen.eval().to(DEV), sd


@pytest.fixture(scope="module")
def golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "hifigan_small.pt"), weights_only=False)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_against_the_reference_golden(golden, mode):
    gen, _ = build(golden["config"], golden["seed"], mode)
    with torch.no_grad():
        wav = gen(golden["mel"].to(DEV))
    assert wav.shape == golden["wav"].shape
    err = float((wav.cpu() - golden["wav"]).abs().max())
    err64 = float((wav.cpu().double() - golden["wav64"]).abs().max())
    print(f"hifigan [{mode}] max |wav - reference fp32| = {err:.2e}, vs fp64 {err64:.2e}")
    assert err < TOL[mode] and err64 < TOL[mode]


def test_ragged_batch_equals_the_per_utterance_reference(golden):
    """one padded launch sequence over utterances of different lengths == the reference called per utterance
    (synthesis/generator.py:160-170): rows past an utterance's end are zeros at every stage, like its own zero padding"""
    gen, _ = build(golden["config"], golden["seed"])
    mels = golden["ragged_mels"]
    lens = [m.shape[0] for m in mels]
    x = torch.zeros(len(mels), 80, max(lens))
    for i, m in enumerate(mels):
        x[i, :, : lens[i]] = m.T
    x[1, :, lens[1]:] = 7.0   # garbage in the padding of the input must not matter either
    with torch.no_grad():
        wav = gen(x.to(DEV), torch.tensor(lens))
    assert wav.shape == (3, 1, max(lens) * 256)
    for i, ref in enumerate(golden["ragged_wavs"]):
        got = wav[i, 0].cpu()
        assert (got[: lens[i] * 256] - ref).abs().max() < 1e-3, i
        assert float(got[lens[i] * 256:].abs().sum()) == 0.0


def test_weight_norm_checkpoint_keys_and_synthesiser(tmp_path, golden):
    """a checkpoint in the bundled file's format ({"generator": {... weight_g / weight_v ...}}) loads through
    Synthesiser exactly like the reference's (hifigan/__init__.py:18-42) and returns int16 samples"""
    cfg = golden["config"]
    ref_gen = hifigan.Generator(hifigan.AttrDict(cfg))
    keys = set(ref_gen.state_dict())
    assert "conv_pre.weight_g" in keys and "ups.3.weight_v" in keys and "resblocks.11.convs2.2.weight_g" in keys
    # weight_g / weight_v of the seeded weights: g = |v| per output slice, v = w
    sd = synthetic.hifigan_state_dict(cfg, seed=golden["seed"])
    ck = {}
    for k, v in sd.items():
        if k.endswith(".weight"):
            ck[k[:-6] + "weight_v"] = v
            ck[k[:-6] + "weight_g"] = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
        else:
            ck[k] = v
    assert set(ck) == keys
    path = os.path.join(tmp_path, "generator_test.pth.tar")
    torch.save({"generator": ck}, path)
    synth = hifigan.Synthesiser(device=DEV, checkpoint=path, config=cfg)
    m = golden["ragged_mels"][0]
    out = synth(m)
    want = (golden["ragged_wavs"][0].numpy() * 32768.0).astype("int16")
    assert out.dtype == np.int16 and out.shape == (1, m.shape[0] * 256)
    assert np.abs(out[0].astype(np.int32) - want.astype(np.int32)).max() <= 33   # 1e-3 of full scale
    many = synth.batch(golden["ragged_mels"])
    for got, ref in zip(many, golden["ragged_wavs"]):
        w = (ref.numpy() * 32768.0).astype("int16")
        assert got.shape == w.shape and np.abs(got.astype(np.int32) - w.astype(np.int32)).max() <= 33


@pytest.mark.parametrize("bsz,t", [(1, 1), (2, 130), (3, 64)])
def test_against_the_oracle_at_other_shapes(bsz, t):
    gen, sd = build(HO.CONFIG, 5)
    g = torch.Generator().manual_seed(t)
    mel = torch.randn(bsz, 80, t, generator=g)
    with torch.no_grad():
        wav = gen(mel.to(DEV))
        ref = HO.generator(sd, mel)
    assert wav.shape == ref.shape == (bsz, 1, t * 256)
    assert (wav.cpu() - ref).abs().max() < 1e-3


def test_no_cpu_path():
    gen = hifigan.Generator(hifigan.AttrDict(HO.CONFIG))
    with pytest.raises(Exception, match="CUDA"):
        gen(torch.zeros(1, 80, 4))


def test_synthesis_stream_vocodes_each_batch_like_the_per_utterance_caller(golden):
    """SynthesisStream(model, vocoder=generator): mel -> int16 waveform on the device for the whole ragged batch ==
    the reference's caller, which vocodes every utterance's own mel (synthesis/generator.py:160-170)"""
    from lightningfastspeech2_b200 import configs
    from lightningfastspeech2_b200.fastspeech2.fastspeech2 import FastSpeech2
    from lightningfastspeech2_b200.pipeline import SynthesisStream

    kw = configs.PRESETS["C2"]
    hp = configs.resolve(kw)
    st = {v: {"min": -3.0, "max": 3.0, "mean": 0.0, "std": 1.0} for v in hp["variances"]}
    model = FastSpeech2(stats=st, phone2id={f"p{i}": i for i in range(80)}, num_workers=0, **kw)
    model.load_state_dict(synthetic.fill_state_dict(model.state_dict(), seed=9))
    model = model.eval().to(DEV)
    gen, _ = build(golden["config"], golden["seed"])
    for compact in (False, True):
        stream = SynthesisStream(model, vocoder=gen, compact=compact)
        batches = [synthetic.make_batch(3, 6, 30, seed=40 + i) for i in range(3)]
        tickets = [stream.submit({k: b[k] for k in ("phones", "speaker")}) for b in batches[:2]]
        results = [stream.collect(t) for t in tickets]
        results.append(stream.collect(stream.submit({k: batches[2][k] for k in ("phones", "speaker")})))
        for res in results:
            assert res["hop"] == 256 and len(res["wav"]) == 3
            for i, n in enumerate(res["lengths"]):
                mel_i = res["mel"][i] if compact else res["mel"][i, :n]
                assert res["wav"][i].dtype == torch.int16 and res["wav"][i].shape == (n * 256,)
                with torch.no_grad():
                    ref = gen(mel_i.T.unsqueeze(0).to(DEV))[0, 0].cpu()
                ref16 = (ref.numpy() * 32768.0).astype("int16")
                diff = np.abs(res["wav"][i].numpy().astype(np.int32) - ref16.astype(np.int32))
                assert diff.max() <= 33, (i, int(diff.max()))     # 1e-3 of full scale: ragged batch vs single utterance
