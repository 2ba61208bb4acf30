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
    "no_epilogue": ["-DLFS2_DIAG_NO_EPILOGUE"],        # accumulators released unread: TMA loads + MMAs only
    "no_epilogue_no_lo_loads": ["-DLFS2_DIAG_NO_EPILOGUE", "-DLFS2_DIAG_NO_LO_LOADS"],
    "ln_no_pass1": ["-DLFS2_DIAG_LN_NO_PASS1"],        # LayerNorm builds: no statistics pass over tensor memory
    "st_no_stencil": ["-DLFS2_DIAG_ST_NO_STENCIL"],    # stencil epilogue: u = z (no z staging, no neighbour reads, 2 barriers less)
    "st_no_stencil_no_stores": ["-DLFS2_DIAG_ST_NO_STENCIL", "-DLFS2_DIAG_NO_STORES"],
    "ln_no_pass1_no_stores": ["-DLFS2_DIAG_LN_NO_PASS1", "-DLFS2_DIAG_NO_STORES"],
}


def build(only=None):
    os.makedirs(AB, exist_ok=True)
    objs = [os.path.join(CSRC, "build", f) for f in os.listdir(os.path.join(CSRC, "build"))
            if f.endswith(".o") and f != "gemm_tc.o"]
    for name, defs in VARIANTS.items():
        if only and name not in only:
            continue
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
    x = ops.split_bf16(torch.randn(64, 2635, 256, generator=g).cuda(), want_f16=True)
    ones, zeros = torch.ones(256, device="cuda"), torch.zeros(256, device="cuda")
    w768 = (torch.randn(768, 256, generator=g) / 16).cuda()
    w256 = (torch.randn(256, 256, generator=g) / 16).cuda()
    b768 = torch.zeros(768, device="cuda")
    dw = ((torch.randn(3, 256, generator=g) / 2).cuda(), zeros)
    head = ((torch.randn(256, generator=g) / 16).cuda(), torch.zeros(1, device="cuda"), None)
    wp768, wp768h, wp256 = ops.split_bf16(w768), ops.split_f16(w768), ops.split_bf16(w256)
    cases = {
        "qkv x3 planes": lambda: ops.gemm_tc(x, wp768, b768, out="planes"),
        "qkv 2-pass f16": lambda: ops.gemm_tc(x, wp768h, b768, out="f16", npass=2),
        "out-proj+res+LN": lambda: ops.gemm_tc(x, wp256, zeros, out="planes", residual=x, gamma=ones, beta=zeros),
        "pred LN+stencil": lambda: ops.predictor_layer_tc(x, wp256, zeros, ones, zeros, next_dw=dw),
        "pred LN+head": lambda: ops.predictor_layer_tc(x, wp256, zeros, ones, zeros, head=head),
    }
    res = []
    for label, fn in cases.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res.append(f"{label} {e0.elapsed_time(e1) / 20:.4f}")
    print(f"{name:24s} " + "   ".join(res) + "  (ms)", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    elif sys.argv[1] == "run":
        for name in (sys.argv[2:] or VARIANTS):
            subprocess.run([sys.executable, os.path.abspath(__file__), "one", name], check=True)
    elif sys.argv[1] == "knobs":  # the shipped library under its environment knobs
        for env in ({}, {"LFS2_GEMM_MULTICAST": "0"}, {"LFS2_GEMM_NTILE": "128"}):
            print(env or "defaults", flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "one", "shipped"], check=True,
                           env=dict(os.environ, **env))
    else:
        one(sys.argv[2])
