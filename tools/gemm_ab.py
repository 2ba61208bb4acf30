"""Timing diagnostics of lfs2_gemm_tc on the C2 decoder QKV shape: which part of the kernel bounds it?

    python tools/gemm_ab.py build   # (here) tools/ab/liblfs2_<variant>.so, variants = -DLFS2_DIAG_* builds (WRONG results)
    python tools/gemm_ab.py run     # (GPU box) times each variant in a fresh process
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
    "no_stores": ["-DLFS2_DIAG_NO_STORES"],            # epilogue stages to shared memory but issues no TMA store
    "no_lo_mmas": ["-DLFS2_DIAG_NO_LO_MMAS"],          # one MMA pass instead of three, all loads kept
    "no_lo_loads": ["-DLFS2_DIAG_NO_LO_LOADS"],        # three MMA passes, only the hi planes are fetched
    "no_lo_at_all": ["-DLFS2_DIAG_NO_LO_MMAS", "-DLFS2_DIAG_NO_LO_LOADS"],
}


def build():
    os.makedirs(AB, exist_ok=True)
    objs = [os.path.join(CSRC, "build", f) for f in os.listdir(os.path.join(CSRC, "build"))
            if f.endswith(".o") and f != "gemm_tc.o"]
    for name, defs in VARIANTS.items():
        obj = os.path.join(AB, f"gemm_tc_{name}.o")
        subprocess.run(["nvcc", *FLAGS, *defs, "-c", os.path.join(CSRC, "gemm_tc.cu"), "-o", obj], check=True)
        subprocess.run(["nvcc", "-shared", "-o", os.path.join(AB, f"liblfs2_{name}.so"), obj, *objs, "-gencode",
                        "arch=compute_100a,code=sm_100a"], check=True)
        print("built", name)


def one(name):
    sys.path.insert(0, ROOT)
    import torch
    from lightningfastspeech2_b200 import _lib
    if name != "shipped":
        _lib.LIB_PATH = os.path.join(AB, f"liblfs2_{name}.so")
    from lightningfastspeech2_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = ops.split_bf16(torch.randn(64, 2635, 256, generator=g).cuda())
    res = []
    for n, kw in ((768, {}), (256, {"ln": True})):
        w = ops.split_bf16((torch.randn(n, 256, generator=g) / 16).cuda())
        b = torch.zeros(n, device="cuda")
        extra = dict(residual=x, gamma=torch.ones(n, device="cuda"), beta=b) if kw else {}
        for _ in range(3):
            ops.gemm_tc(x, w, b, out="planes", **extra)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.gemm_tc(x, w, b, out="planes", **extra)
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 20)
    print(f"{name:14s} qkv (n=768) {res[0]:.4f} ms   out-proj + residual + LN (n=256) {res[1]:.4f} ms", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    elif sys.argv[1] == "run":
        for name in VARIANTS:
            subprocess.run([sys.executable, os.path.abspath(__file__), "one", name], check=True)
    elif sys.argv[1] == "knobs":  # the shipped library under its environment knobs
        for env in ({}, {"LFS2_GEMM_MULTICAST": "0"}, {"LFS2_GEMM_NTILE": "128"}):
            print(env or "defaults", flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "one", "shipped"], check=True,
                           env=dict(os.environ, **env))
    else:
        one(sys.argv[2])
