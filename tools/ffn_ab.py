"""Timing diagnostics of lfs2_ffn_fused_tc on the C2 decoder shape (168640 x 256, F = 1024): what bounds the kernel?

    python tools/ffn_ab.py build   # (here) tools/ab/liblfs2_ffn_<variant>.so, variants = -DLFS2_FFN_DIAG_* builds (WRONG results)
    python tools/ffn_ab.py run     # (GPU box) times each variant in a fresh process, npass 3 and 2
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
    "baseline": [],
    "no_e1": ["-DLFS2_FFN_DIAG_NO_E1"],               # the intermediate is not converted (acc1 bits reused as v)
    "no_ln": ["-DLFS2_FFN_DIAG_NO_LN"],               # no LayerNorm epilogue, no stores
    "no_stores": ["-DLFS2_FFN_DIAG_NO_STORES"],       # LayerNorm epilogue without its hi/lo TMA stores
    "no_e1_no_ln": ["-DLFS2_FFN_DIAG_NO_E1", "-DLFS2_FFN_DIAG_NO_LN"],   # TMA + MMA only
    "no_wlo_loads": ["-DLFS2_FFN_DIAG_NO_WLO_LOADS"],   # all MMAs, but the lo weight planes are not fetched (-1 MB of 2.6 per tile)
    "no_wlo_no_ln": ["-DLFS2_FFN_DIAG_NO_WLO_LOADS", "-DLFS2_FFN_DIAG_NO_LN"],
    "timeline": ["-DLFS2_FFN_TIMELINE"],              # correct results + clock64 stamps per CTA and tile (see `timeline`)
}


def build():
    os.makedirs(AB, exist_ok=True)
    objs = [os.path.join(CSRC, "build", f) for f in os.listdir(os.path.join(CSRC, "build"))
            if f.endswith(".o") and f != "ffn_fused_tc.o"]
    for name, defs in VARIANTS.items():
        obj = os.path.join(AB, f"ffn_{name}.o")
        subprocess.run(["nvcc", *FLAGS, *defs, "-c", os.path.join(CSRC, "ffn_fused_tc.cu"), "-o", obj], check=True)
        subprocess.run(["nvcc", "-shared", "-o", os.path.join(AB, f"liblfs2_ffn_{name}.so"), obj, *objs, "-gencode",
                        "arch=compute_100a,code=sm_100a"], check=True)
        os.remove(obj)
        print("built", name)


def one(name):
    sys.path.insert(0, ROOT)
    import torch
    from lightningfastspeech2_b200 import _lib
    if name != "shipped":
        _lib.LIB_PATH = os.path.join(AB, f"liblfs2_ffn_{name}.so")
    from lightningfastspeech2_b200 import ops
    g = torch.Generator().manual_seed(0)
    m, f, d = 64 * 2635, 1024, 256
    x = torch.randn(m, d, generator=g).cuda()
    xp = ops.split_bf16(x, want_f16=True)
    w1, w2 = (torch.randn(f, d, generator=g) / 16).cuda(), (torch.randn(d, f, generator=g) / 32).cuda()
    b1, b2, gam = torch.zeros(f, device="cuda"), torch.zeros(d, device="cuda"), torch.ones(d, device="cuda")
    res = []
    for npass in (3, 2):
        sp = ops.split_f16 if npass == 2 else ops.split_bf16
        u = ops.Planes(xp.h, None) if npass == 2 else xp
        args = (u, sp(w1), b1, sp(w2), b2, xp, gam, b2)
        for _ in range(3):
            ops.ffn_fused_tc(*args, npass=npass, want_f16=npass == 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.ffn_fused_tc(*args, npass=npass, want_f16=npass == 2)
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 20)
    tf = [4.0 * m * d * f * n / (t * 1e-3) / 1e12 for n, t in zip((3, 2), res)]
    print(f"{name:14s} npass=3 {res[0]:.4f} ms ({tf[0]:.0f} TF/s issued)   npass=2 {res[1]:.4f} ms ({tf[1]:.0f} TF/s issued)", flush=True)


def timeline():
    """per-tile stamps of CTA 0..3 (cycles relative to the tile's first stamp), npass = 2 then 3"""
    import ctypes
    sys.path.insert(0, ROOT)
    import torch
    from lightningfastspeech2_b200 import _lib
    _lib.LIB_PATH = os.path.join(AB, "liblfs2_ffn_timeline.so")
    from lightningfastspeech2_b200 import ops
    g = torch.Generator().manual_seed(0)
    m, f, d = 64 * 2635, 1024, 256
    x = torch.randn(m, d, generator=g).cuda()
    xp = ops.split_bf16(x, want_f16=True)
    w1, w2 = (torch.randn(f, d, generator=g) / 16).cuda(), (torch.randn(d, f, generator=g) / 32).cuda()
    b1, b2, gam = torch.zeros(f, device="cuda"), torch.zeros(d, device="cuda"), torch.ones(d, device="cuda")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = ["epi:tile start", "epi:E1 done", "epi:acc2_full seen", "epi:stats done", "epi:LN+stores done",
             "mma:tile start", "mma:first G2 may go", "mma:residual issue"]
    for npass in (2, 3):
        sp = ops.split_f16 if npass == 2 else ops.split_bf16
        u = ops.Planes(xp.h, None) if npass == 2 else xp
        for _ in range(3):
            ops.ffn_fused_tc(u, sp(w1), b1, sp(w2), b2, xp, gam, b2, npass=npass, want_f16=npass == 2)
        torch.cuda.synchronize()
        buf = (ctypes.c_longlong * (148 * 16 * 8))()
        assert lib.lfs2_ffn_timeline(buf) == 0
        t = torch.tensor(list(buf), dtype=torch.int64).view(148, 16, 8)
        print(f"npass={npass}: stamps in cycles relative to the epilogue's start of tile 1 of each CTA (CTA 0, 1, 74)")
        for cta in (0, 1, 74):
            base = int(t[cta, 1, 0])
            for it in range(1, 5):
                row = "  ".join(f"{names[k].split(':')[0][0]}{k}={int(t[cta, it, k]) - base:7d}" for k in (5, 0, 6, 1, 7, 2, 3, 4))
                print(f"  cta {cta:3d} tile {it}: {row}")
        per_tile = (t[:, 2:8, 0] - t[:, 1:7, 0]).double().mean()
        ln = (t[:, 1:8, 4] - t[:, 1:8, 2]).double().mean()
        st = (t[:, 1:8, 3] - t[:, 1:8, 2]).double().mean()
        wait_acc2 = (t[:, 1:8, 2] - t[:, 1:8, 1]).double().mean()
        g2_wait = (t[:, 2:8, 6] - t[:, 2:8, 5]).double().mean()
        print(f"  mean cycles: tile period {per_tile:.0f}, E1 loop {float((t[:, 1:8, 1] - t[:, 1:8, 0]).double().mean()):.0f}, "
              f"wait for acc2_full {wait_acc2:.0f}, LN pass 1 {st:.0f}, LN total {ln:.0f}, "
              f"mma: tile start -> first G2 issued {g2_wait:.0f}")
    print("legend:", ", ".join(f"{k}={n}" for k, n in enumerate(names)))


if __name__ == "__main__":
    if sys.argv[1] == "timeline":
        timeline()
    elif sys.argv[1] == "build":
        build()
    elif sys.argv[1] == "run":
        for env in ({}, {"LFS2_FFN_MULTICAST": "0"}):
            print(env or "defaults (2-CTA multicast)", flush=True)
            for name in (sys.argv[2:] or VARIANTS):
                subprocess.run([sys.executable, os.path.abspath(__file__), "one", name], check=True,
                               env=dict(os.environ, **env))
    else:
        one(sys.argv[2])
