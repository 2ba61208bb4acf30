"""A/B timing of lfs2_attention_tc variants on the C2 decoder shape.

    python tools/attn_ab.py build      # (here) builds tools/ab/liblfs2_<name>.so: HEAD's kernel + the tuning-knob variants
    python tools/attn_ab.py run        # (GPU box) times every variant, fp32-parity and bf16 mode, in fresh processes
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AB = os.path.join(ROOT, "tools", "ab")
CSRC = os.path.join(ROOT, "lightningfastspeech2_b200", "csrc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]
VARIANTS = {
    "head": None,  # attention_tc.cu as committed at HEAD
    "default": [],  # the source defaults (two 2-tile S/P buffers in both modes)
    "g1b3": ["-DLFS2_ATTN_G3=1", "-DLFS2_ATTN_BUFS3=3", "-DLFS2_ATTN_G1=1", "-DLFS2_ATTN_BUFS1=3"],  # three 1-tile buffers
    "g1b2": ["-DLFS2_ATTN_G3=1", "-DLFS2_ATTN_BUFS3=2", "-DLFS2_ATTN_G1=1", "-DLFS2_ATTN_BUFS1=2"],
}


def build():
    os.makedirs(AB, exist_ok=True)
    objs = [os.path.join(CSRC, "build", f) for f in os.listdir(os.path.join(CSRC, "build"))
            if f.endswith(".o") and f != "attention_tc.o"]
    for name, defs in VARIANTS.items():
        src = os.path.join(CSRC, "attention_tc.cu")
        if defs is None:
            src = os.path.join(CSRC, "_attention_tc_head.cu")
            with open(src, "w") as f:
                f.write(subprocess.run(["git", "show", "HEAD:lightningfastspeech2_b200/csrc/attention_tc.cu"], cwd=ROOT,
                                       capture_output=True, text=True, check=True).stdout)
        obj = os.path.join(AB, f"attention_tc_{name}.o")
        subprocess.run(["nvcc", *FLAGS, *(defs or []), "-c", src, "-o", obj], check=True)
        subprocess.run(["nvcc", "-shared", "-o", os.path.join(AB, f"liblfs2_{name}.so"), obj, *objs, "-gencode",
                        "arch=compute_100a,code=sm_100a"], check=True)
        if defs is None:
            os.remove(src)
        print("built", name)


def one(name):
    sys.path.insert(0, ROOT)
    import torch
    from lightningfastspeech2_b200 import _lib
    _lib.LIB_PATH = os.path.join(AB, f"liblfs2_{name}.so")
    from lightningfastspeech2_b200 import ops, synthetic
    import bench
    b = synthetic.make_batch(bench.BATCH, bench.MIN_LEN, bench.MAX_LEN, seed=2)
    g = torch.Generator().manual_seed(0)
    lens = ((b["phones"] != 0).sum(1) * 5.2).long().clamp(max=2635)  # frames per utterance, like the bench batch
    t = int(lens.max())
    kpm = (torch.arange(t)[None] >= lens[:, None]).cuda()
    qkv = ops.split_bf16(torch.randn(bench.BATCH, t, 768, generator=g).cuda())
    out = []
    for npass in (3, 1):
        for _ in range(3):
            ops.attention_tc(qkv, kpm, 2, npass=npass)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.attention_tc(qkv, kpm, 2, npass=npass)
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) / 20)
    print(f"{name:16s} fp32 {out[0]:.4f} ms  bf16 {out[1]:.4f} ms", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    elif sys.argv[1] == "run":
        for rep in range(2):
            for name in VARIANTS:
                subprocess.run([sys.executable, os.path.abspath(__file__), "one", name], check=True)
    else:
        one(sys.argv[2])
