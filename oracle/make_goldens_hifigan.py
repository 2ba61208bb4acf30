#!/usr/bin/env python
"""Golden vectors of the HiFi-GAN generator from the UNMODIFIED reference (run in the authoring container only):

    python oracle/make_goldens_hifigan.py      # writes tests/golden/hifigan_small.pt

The reference's ``Generator`` (litfass/third_party/hifigan/models.py, loaded from its file so that none of litfass'
audio dependencies is imported) is built from the reference's config.json, given the seeded weights of
``synthetic.hifigan_state_dict`` (weight_norm removed, as Synthesiser does at __init__.py:31) and run on CPU in fp32 and
-- as a ``.double()`` copy -- in fp64 on seeded mels: one plain batch and the utterances of a ragged batch one by one
(what SpeechGenerator.generate_samples does, synthesis/generator.py:160-170).  Only the small inputs/outputs are stored;
the weights are regenerated from the seed wherever the golden is replayed.  When the bundled generator_universal.pth.tar
is present, its output on the first golden mel is stored too (a 9 k-sample vector), so the oracle restatement can be
checked against real trained weights here.
"""
import copy
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lightningfastspeech2_b200 import synthetic  # noqa: E402

REF = "/root/reference/litfass/third_party/hifigan"


class AttrDict(dict):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.__dict__ = self


def reference_generator():
    spec = importlib.util.spec_from_file_location("ref_hifigan_models", os.path.join(REF, "models.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    h = AttrDict(json.load(open(os.path.join(REF, "config.json"))))
    return mod.Generator(h), h


def main():
    seed = 3
    gen, h = reference_generator()
    gen.eval()
    gen.remove_weight_norm()
    sd = synthetic.hifigan_state_dict(h, seed=seed)
    gen.load_state_dict(sd, strict=True)
    gen64 = copy.deepcopy(gen).double()
    g = torch.Generator().manual_seed(11)
    mel = torch.randn(2, 80, 23, generator=g)
    lens = [19, 7, 31]
    ragged = [torch.randn(n, 80, generator=g) for n in lens]
    with torch.no_grad():
        out = gen(mel)
        out64 = gen64(mel.double())
        per_utt = [gen(m.T.unsqueeze(0))[0, 0] for m in ragged]
        per_utt64 = [gen64(m.T.unsqueeze(0).double())[0, 0] for m in ragged]
    golden = {"seed": seed, "config": dict(h), "mel": mel, "wav": out, "wav64": out64, "ragged_mels": ragged,
              "ragged_wavs": per_utt, "ragged_wavs64": per_utt64, "state_dict_keys": sorted(gen.state_dict())}
    ck = os.path.join(REF, "generator_universal.pth.tar")
    if os.path.exists(ck):
        g2, _ = reference_generator()
        g2.load_state_dict(torch.load(ck, map_location="cpu", weights_only=False)["generator"])
        g2.eval()
        g2.remove_weight_norm()
        with torch.no_grad():
            golden["universal_wav_first_mel"] = g2(mel[:1] * 2 - 4)[0, 0]
    path = os.path.join(ROOT, "tests", "golden", "hifigan_small.pt")
    torch.save(golden, path)
    print("wrote", path, os.path.getsize(path), "bytes; |wav| mean", float(out.abs().mean()), "max", float(out.abs().max()))


if __name__ == "__main__":
    main()
