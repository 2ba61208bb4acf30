"""clock64 timeline of attention_tc_pp_kernel on the C2 decoder shape (B=64, T=2635, d=256, 2 heads, fp16 operands):

    python tools/attn_timeline.py build   # (here) tools/ab/liblfs2_attn_timeline.so (-DLFS2_ATTN_TIMELINE)
    python tools/attn_timeline.py run     # (GPU box) per-CTA stamps -> where a CTA's life goes
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AB = os.path.join(ROOT, "tools", "ab")
CSRC = os.path.join(ROOT, "lightningfastspeech2_b200", "csrc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def build():
    os.makedirs(AB, exist_ok=True)
    objs = [os.path.join(CSRC, "build", f) for f in os.listdir(os.path.join(CSRC, "build"))
            if f.endswith(".o") and f != "attention_tc_pp.o"]
    obj = os.path.join(AB, "attn_timeline.o")
    subprocess.run(["nvcc", *FLAGS, "-DLFS2_ATTN_TIMELINE", "-c", os.path.join(CSRC, "attention_tc_pp.cu"), "-o", obj], check=True)
    subprocess.run(["nvcc", "-shared", "-o", os.path.join(AB, "liblfs2_attn_timeline.so"), obj, *objs, "-gencode",
                    "arch=compute_100a,code=sm_100a"], check=True)
    os.remove(obj)
    print("built")


def run():
    sys.path.insert(0, ROOT)
    import torch
    from lightningfastspeech2_b200 import _lib
    _lib.LIB_PATH = os.path.join(AB, "liblfs2_attn_timeline.so")
    from lightningfastspeech2_b200 import ops, synthetic
    b, t, d, nh = 64, 2635, 256, 2
    g = torch.Generator().manual_seed(0)
    qkv = ops.Planes((torch.randn(b, t, 3 * d, generator=g)).half().cuda(), None)
    # the bench batch's frame counts (seed 2): realistic key-padding masks
    lens = torch.randint(300, t + 1, (b,), generator=g)
    lens[0] = t
    kpm = (torch.arange(t)[None, :] >= lens[:, None]).cuda()
    for _ in range(3):
        ops.attention_tc(qkv, kpm, nh, npass=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.attention_tc(qkv, kpm, nh, npass=1)
    e1.record()
    torch.cuda.synchronize()
    print(f"kernel: {e0.elapsed_time(e1) / 10:.4f} ms per launch")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    n = 4096 * 12
    buf = (ctypes.c_longlong * n)()
    assert lib.lfs2_attn_timeline(buf) == 0
    tl = torch.tensor(list(buf), dtype=torch.int64).view(4096, 12).double()
    nct = min(4096, ((t + 127) // 128) * nh * b)
    tl = tl[:nct]
    live = tl[:, 0] > 0
    x = tl[live]
    nt = x[:, 0]
    names = ["tiles", "Q landed", "mma: wait P (sum)", "mma: wait V (sum)", "mma: last PV issued", "softmax: first S seen",
             "softmax: wait S (sum)", "softmax: last P written", "O complete", "O rows stored"]
    print(f"{int(live.sum())} CTAs with work of the first {nct}; mean key tiles per CTA {nt.mean():.1f}")
    for k in range(1, 10):
        print(f"  {names[k]:28s} mean {x[:, k].mean():9.0f} cycles   per tile {(x[:, k] / nt).mean():7.0f}")
    life = x[:, 9]
    print(f"  CTA life per key tile: {(life / nt).mean():.0f} cycles; prologue (to first S) {x[:, 5].mean():.0f}; "
          f"epilogue (last P -> rows stored) {(x[:, 9] - x[:, 7]).mean():.0f}")
    print(f"  per tile: softmax busy {((x[:, 7] - x[:, 5] - x[:, 6]) / nt).mean():.0f}, softmax waiting for S {(x[:, 6] / nt).mean():.0f}, "
          f"mma waiting for P {(x[:, 2] / nt).mean():.0f}, mma waiting for V {(x[:, 3] / nt).mean():.0f}")
    # concurrency: how many CTAs were alive on SM 0 .. over time is not reconstructed; report slots instead
    sm = x[:, 10].long()
    per_sm = torch.bincount(sm, minlength=148)
    print(f"  CTAs per SM (of the first {nct}): min {int(per_sm.min())} max {int(per_sm.max())}")


if __name__ == "__main__":
    build() if sys.argv[1] == "build" else run()
