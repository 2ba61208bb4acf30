#!/usr/bin/env python
"""bench.py -- synthesis throughput of the B200-native mel-generation path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl lfs2|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "C2"): LightSpeech depthwise-separable FastSpeech2
(reference constructor defaults, 2 frame-level variances, 7.4 M params), batch of 64
synthetic utterances with 32..512 phonemes, `model(batch, inference=True)`.
One step = one forward of the whole hot path (Encoder -> VarianceAdaptor incl.
LengthRegulator -> Decoder -> mel Linear) over the batch.  Metric = valid mel frames
((~tgt_mask).sum()) per second, whole job (all ranks).  N>1 = one process per GPU, each
rank synthesising its own 64-utterance shard (weak scaling, no data-path collective).

Keys beyond the base contract: `roofline` (dominant kernel, CUDA-event timed inside this
script), `cpu_baseline` (the oracle port on host cores, bounded sample), `e2e` (same
metric through the public module API with HOST inputs: pinned H2D + D2H of the mel inside
the timed region).  `--impl reference` times the CPU port (oracle/) of the reference path.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from lightningfastspeech2_b200 import configs, synthetic  # noqa: E402

METRIC = "valid mel-frames/sec (synthesis)"
UNIT = "mel-frames/s"
PRESET = "C2"
BATCH, MIN_LEN, MAX_LEN = 64, 32, 512
WORKLOAD = (f"C2 LightSpeech depthwise FastSpeech2 (7.4M params) synthesis, batch={BATCH} utterances/GPU, "
            f"phoneme len U[{MIN_LEN},{MAX_LEN}], fp32-parity mode")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm": d["hbm_gbs"], "tensor": d["bf16_tflops"], "tensor_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm": 6650.0, "tensor": 1590.0, "tensor_sustained": 1400.0, "src": "fallback"}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed
    `ncu --set full` capture of this workload (profiles/ncu_traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = f"/tmp/lfs2_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            top = sorted(sm)[len(sm) // 2:]  # samples under load = upper half
            out = {"sm_mhz": statistics.median(top), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def dist_env():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    return rank, world, local


def build_model(device, preset=PRESET):
    from lightningfastspeech2_b200.fastspeech2.fastspeech2 import FastSpeech2

    kw = configs.PRESETS[preset]
    hp = configs.resolve(kw)
    stats = {v: {"min": -3.0, "max": 3.0, "mean": 0.0, "std": 1.0} for v in hp["variances"]}
    model = FastSpeech2(stats=stats, phone2id={f"p{i}": i for i in range(80)}, num_workers=0, **kw)
    sd = synthetic.fill_state_dict(model.state_dict(), seed=0)
    model.load_state_dict(sd)
    hp["stats"] = stats
    return model.eval().to(device), sd, hp


def cpu_port_throughput(sd, hp, batch, nutt, repeats=1):
    """Oracle port (torch CPU ops = what the reference's nn.Modules dispatch to) on the first
    `nutt` utterances of the batch, all host threads.  Returns (frames/s, seconds, frames)."""
    from oracle import fs2_oracle as O

    torch.set_num_threads(os.cpu_count())
    sub = {"phones": batch["phones"][:nutt].contiguous(), "speaker": batch["speaker"][:nutt].contiguous()}
    keep = int((sub["phones"] != 0).sum(1).max())
    sub["phones"] = sub["phones"][:, :keep].contiguous()
    best = None
    frames = 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        with torch.no_grad():
            r = O.forward(sd, hp, sub, inference=True)
        dt = time.perf_counter() - t0
        frames = int((~r["tgt_mask"]).sum())
        best = dt if best is None else min(best, dt)
    return frames / best, best, frames


# ---------------------------------------------------------------------------------------------
# train step (BASELINE.json configs[3], "C4"): forward + FastSpeech2Loss + backward + gradient
# all-reduce + AdamW/Noam on the 76 M model, global batch 64 split across the ranks (strong scaling)
TRAIN_PRESET = "C4"
TRAIN_BATCH, TRAIN_MIN_LEN, TRAIN_MAX_LEN = 64, 32, 256
TRAIN_WORKLOAD = (f"C4 train step (forward + loss + backward + grad all-reduce + AdamW/Noam), 76M-parameter "
                  f"depthwise FastSpeech2 (d=768, 4 enc + 5 dec FFTBlocks, 3 variances), global batch={TRAIN_BATCH} "
                  f"utterances, phoneme len U[{TRAIN_MIN_LEN},{TRAIN_MAX_LEN}], durations U[1,9], reference-default dropout "
                  f"(0.1 / 0.5) in the CUDA arm, none in the CPU port")


def train_batch(hp, rank, world):
    """The rank's shard of the global train batch: utterances sorted by length and dealt in snake
    order, each rank padded to its own maximum length."""
    from lightningfastspeech2_b200.sharding import shard_batch

    full = synthetic.make_batch(TRAIN_BATCH, TRAIN_MIN_LEN, TRAIN_MAX_LEN, seed=4)
    return synthetic.add_train_targets(shard_batch(full, rank, world), hp["variances"], seed=4 + rank)


def build_train_model(device, preset=TRAIN_PRESET):
    from lightningfastspeech2_b200.fastspeech2.fastspeech2 import FastSpeech2

    kw = configs.PRESETS[preset]
    hp = configs.resolve(kw)
    stats = {v: {"min": -3.0, "max": 3.0, "mean": 0.0, "std": 1.0} for v in hp["variances"]}
    model = FastSpeech2(stats=stats, phone2id={f"p{i}": i for i in range(80)}, num_workers=0, **kw)
    sd = synthetic.fill_state_dict(model.state_dict(), seed=0)
    model.load_state_dict(sd)
    hp["stats"] = stats
    if device is not None:
        model = model.to(device).train()
    return model, sd, hp


def cpu_port_train_step(sd, hp, batch, nutt):
    """Oracle port of one train step (forward + loss + autograd backward + AdamW/Noam update of every
    parameter) on the first `nutt` utterances, all host threads.  Returns seconds."""
    from oracle import fs2_oracle as O

    torch.set_num_threads(os.cpu_count())
    sub = {k: (v[:nutt].contiguous() if torch.is_tensor(v) else v) for k, v in batch.items()}
    keep = int((sub["phones"] != 0).sum(1).max())
    tm = int(sub["duration"][:, :keep].sum(1).max())
    sub["phones"], sub["duration"] = sub["phones"][:, :keep].contiguous(), sub["duration"][:, :keep].contiguous()
    sub["mel"] = sub["mel"][:, :tm].contiguous()
    for v in hp["variances"]:
        sub[f"variances_{v}"] = sub[f"variances_{v}"][:, :tm].contiguous()
    t0 = time.perf_counter()
    _, grads = O.gradients(sd, hp, sub)
    for k, g in grads.items():
        O.adamw_noam_step(sd[k], g, torch.zeros_like(g), torch.zeros_like(g), 1, hp["lr"], hp["warmup_steps"])
    return time.perf_counter() - t0, int(sub["duration"].sum())


def run_train_steps(model, batch, opt, sch, steps, world):
    for _ in range(steps):
        loss = model.training_step(batch, 0)
        loss.backward()
        opt.grad_scale = 1.0 / model.allreduce_gradients()
        opt.step()
        sch.step()


def measure_train(args, dev, rank, world, barrier):
    """-> dict for the "train" key of the JSON line (rank 0) / None"""
    import torch.distributed as dist

    from lightningfastspeech2_b200 import _lib, ops

    model, sd, hp = build_train_model(dev)
    model.set_compute_mode(args.train_mode)
    model.log_losses = False
    batch = train_batch(hp, rank, world)
    dbatch = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    (opt,), (schd,) = model.configure_optimizers()
    sch = schd["scheduler"]
    run_train_steps(model, dbatch, opt, sch, 3, world)
    barrier()
    calls0 = _lib.CALLS
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_train_steps(model, dbatch, opt, sch, args.train_steps, world)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.train_steps
    launches = (_lib.CALLS - calls0) // args.train_steps
    ops.PROFILE = {}
    run_train_steps(model, dbatch, opt, sch, 1, world)
    prof = ops.collect_profile()
    ops.PROFILE = None
    loss_vals = model.loss.last_buffer.tolist()
    frames = int(batch["duration"].sum())

    def timed_variant(steps):
        run_train_steps(model, dbatch, opt, sch, 2, world)
        barrier()
        c0 = _lib.CALLS
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run_train_steps(model, dbatch, opt, sch, steps, world)
        b.record()
        barrier()
        t = a.elapsed_time(b) / steps
        if world > 1:
            tt = torch.tensor([t], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt[0])
        return t, (_lib.CALLS - c0) // steps

    # the same step as length-sorted sub-batches (model.train_length_buckets: same losses and gradients, PAD rows beyond
    # each bucket's longest utterance + conv halo never computed) and in the single-pass bf16 mode
    variants = {"length_buckets": [], "bf16_mode": None}
    for nb in args.train_buckets:
        if nb > 1 and int(batch["phones"].shape[0]) >= 2 * nb:
            model.train_length_buckets = nb
            t, c = timed_variant(args.train_steps)
            variants["length_buckets"].append({"train_length_buckets": nb, "ms_per_step": t, "gpu_launches_per_step": c,
                                               "total_loss": float(model.loss.last_buffer[-1])})
    model.train_length_buckets = 1
    if args.train_mode != "bf16":
        model.set_compute_mode("bf16")
        t, c = timed_variant(args.train_steps)
        variants["bf16_mode"] = {"ms_per_step": t, "gpu_launches_per_step": c,
                                 "total_loss": float(model.loss.last_buffer[-1])}
        best = min(variants["length_buckets"], key=lambda r: r["ms_per_step"], default=None)
        if best is not None:
            model.train_length_buckets = best["train_length_buckets"]
            t, c = timed_variant(args.train_steps)
            variants["bf16_mode"]["with_length_buckets"] = {"train_length_buckets": best["train_length_buckets"],
                                                            "ms_per_step": t, "gpu_launches_per_step": c}
            model.train_length_buckets = 1
        model.set_compute_mode(args.train_mode)
    nparams = sum(p.numel() for p in model.parameters() if p.requires_grad)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        c = torch.tensor([frames], device=dev, dtype=torch.int64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        frames = int(c[0])
    del model, opt
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    total_ms = sum(v["ms"] for v in prof.values()) or 1.0
    out = {"metric": "ms/step (train)", "ms_per_step": ms, "steps": args.train_steps, "warmup": 3,
           "scaling": "strong", "n_gpus": world, "dtype": {"simt": "f32", "fp32": "f32", "bf16": "bf16"}[args.train_mode],
           "compute_mode": args.train_mode, "valid_frames_per_step": frames, "frames_per_s": frames / (ms * 1e-3),
           "parameters": nparams, "gpu_launches_per_step": launches,
           "allreduce_bytes_per_step": 0 if world == 1 else 4 * nparams,
           "final_losses": {"total": loss_vals[-1]},
           "length_buckets": {"what": "the same step with model.train_length_buckets = n: n length-sorted sub-batches, each "
                                      "padded to its own longest utterance + the conv halo; losses and gradients equal the "
                                      "one-tensor step up to fp32 summation order (tests/test_gpu_round2.py)",
                              "runs": variants["length_buckets"]},
           "bf16_mode": dict(variants["bf16_mode"] or {}, what="the same step with model.set_compute_mode('bf16'): single-pass "
                             "bf16 MMA operands, fp32 accumulation / LayerNorm / softmax / master weights / optimizer (the "
                             "reference's recipe trains with --precision 16, scripts/train.sh)"),
           "config": {"workload": TRAIN_WORKLOAD, "preset": TRAIN_PRESET, "utterances_rank0": int(batch["phones"].shape[0]),
                      "padded_phones_rank0": int(batch["phones"].shape[1]), "mel_frames_rank0": int(batch["mel"].shape[1]),
                      "parallelism": f"data-parallel x{world}, one NCCL all-reduce of the flat fp32 gradient buffer"},
           "kernel_shares": {k: round(v["ms"] / total_ms, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:12]}}
    if world == 1 and args.train_cpu_utts > 0:
        nutt = args.train_cpu_utts  # grow the sample until one oracle train step takes ~8 s of CPU work
        while True:
            secs, cfr = cpu_port_train_step(sd, hp, batch, nutt)
            if secs >= 8.0 or nutt >= 32:
                break
            nutt = min(32, nutt * (4 if secs < 1.5 else 2))
        out["cpu_baseline"] = {"value": secs * 1e3 * frames / max(cfr, 1), "unit": "ms/step (extrapolated by frames)",
                               "cores": os.cpu_count(), "kind": "port",
                               "sample": f"first {nutt} utterances of the batch ({cfr} frames), one oracle train step "
                                         f"(forward + loss + autograd backward + AdamW/Noam) of {secs:.1f} s, scaled to "
                                         f"{frames} frames"}
    return out


# ---------------------------------------------------------------------------------------------
# secondary sections (rank 0 only, local synchronisation only)
def parity_vs_oracle(model, sd, hp, host_batch, nutt, modes):
    """max |mel_cuda - mel_oracle| on a slice of the TIMED batch: the first `nutt` utterances, padded to the full
    batch's phoneme length exactly as in the timed tensors (padding leaks through the FFN convolutions, so the padded
    length is part of the input), oracle (CPU port of the reference) vs this model in each compute mode.  The oracle's
    discrete decisions (rounded durations, bucket indices) are forced on the CUDA run when any of them flipped
    (SURVEY 0.6); flips are reported.  modes: {label: (compute_mode, skip_pad_rows)}."""
    from oracle import fs2_oracle as O

    sub = {"phones": host_batch["phones"][:nutt].contiguous(), "speaker": host_batch["speaker"][:nutt].contiguous()}
    torch.set_num_threads(os.cpu_count())
    t0 = time.perf_counter()
    with torch.no_grad():
        ref = O.forward(sd, hp, sub, inference=True)
    out = {"utterances": nutt, "padded_phones": int(sub["phones"].shape[1]), "mel_shape": list(ref["mel"].shape),
           "oracle_seconds": round(time.perf_counter() - t0, 2), "modes": {}}
    valid = ~ref["tgt_mask"]
    force = {"duration_rounded": ref["duration_rounded"], "bucket_idx": {v: ref[f"_bucket_{v}"] for v in hp["variances"]},
             "want_idx": True}
    old_mode, old_skip = model.compute_mode, model.skip_pad_rows
    try:
        for label, (mode, skip) in modes.items():
            model.set_compute_mode(mode)
            model.skip_pad_rows = skip
            with torch.no_grad():
                r = model(sub, inference=True, force={"want_idx": True})
            flips = int((r["duration_rounded"].cpu() != ref["duration_rounded"]).sum())
            if r["mel"].shape == ref["mel"].shape:
                flips += sum(int((r[f"_bucket_{v}"].cpu() != ref[f"_bucket_{v}"]).sum()) for v in hp["variances"])
            if flips or r["mel"].shape != ref["mel"].shape:
                with torch.no_grad():
                    r = model(sub, inference=True, force=force)
            d = (r["mel"].cpu() - ref["mel"]).abs()
            out["modes"][label] = {"max_abs_mel_err_valid_frames": float(d[valid].max()),
                                   "max_abs_mel_err_all_positions": None if skip else float(d.max()),
                                   "masks_equal": bool(torch.equal(r["tgt_mask"].cpu(), ref["tgt_mask"])),
                                   "discrete_decision_flips_before_forcing": flips,
                                   "tolerance": 1e-2 if mode == "bf16" else 1e-3}
    finally:
        model.set_compute_mode(old_mode)
        model.skip_pad_rows = old_skip
    return out


def measure_c1(dev, steps=30):
    """BASELINE.json configs[0] ("C1"): ming024-like dense-conv FastSpeech2 (k = 9, 24.5 M params), ONE 128-phoneme
    utterance: latency of model(batch, inference=True) incl. the H2D of the inputs and a D2H of the mel, next to the
    CPU port on the same input."""
    from oracle import fs2_oracle as O

    model, sd, hp = build_model(dev, preset="C1")
    batch = synthetic.make_batch(1, 128, 128, seed=1234)
    pinned = {k: v.pin_memory() for k, v in batch.items() if k in ("phones", "speaker")}

    def once():
        with torch.no_grad():
            r = model(pinned, inference=True)
        return r["mel"].cpu(), r

    for _ in range(5):
        once()
    lat = []
    for _ in range(steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mel, r = once()
        lat.append((time.perf_counter() - t0) * 1e3)
    frames = int((~r["tgt_mask"]).sum())
    # the same call with model.cuda_graphs = True: the ~100 short launches of a 1-utterance call replayed as two CUDA
    # graphs (encoder side | one host read-back of the frame count | decoder side); bit-identical results
    model.cuda_graphs = True
    for _ in range(4):
        mel_g, _ = once()
    lat_g = []
    for _ in range(steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mel_g, _ = once()
        lat_g.append((time.perf_counter() - t0) * 1e3)
    graphs_identical = bool(torch.equal(mel_g, mel))
    model.cuda_graphs = False
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    resident = {k: v.to(dev) for k, v in pinned.items()}
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        with torch.no_grad():
            model(resident, inference=True)
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / steps
    torch.set_num_threads(os.cpu_count())
    cpu = []
    for _ in range(4):
        t0 = time.perf_counter()
        with torch.no_grad():
            ref = O.forward(sd, hp, batch, inference=True)
        cpu.append((time.perf_counter() - t0) * 1e3)
    cpu_ms = statistics.median(cpu[1:])
    err = None
    if ref["mel"].shape == mel.shape:
        err = float((mel - ref["mel"]).abs().max())
    ms = statistics.median(lat)
    return {"workload": "C1: dense-conv FastSpeech2 (k=9, d=256, 4+4 FFTBlocks, 24.5M params), 1 utterance x 128 phonemes, "
                        "fp32-parity mode (BASELINE.json configs[0])",
            "latency_ms_e2e_median": ms, "latency_ms_device": dev_ms, "valid_frames": frames,
            "latency_ms_e2e_median_cuda_graphs": statistics.median(lat_g),
            "cuda_graphs_bit_identical": graphs_identical,
            "value": frames / (ms * 1e-3), "unit": UNIT, "steps": steps,
            "cpu_baseline": {"latency_ms": cpu_ms, "value": int((~ref["tgt_mask"]).sum()) / (cpu_ms * 1e-3), "unit": UNIT,
                             "cores": os.cpu_count(), "kind": "port", "sample": "the same utterance, median of 3 runs"},
            "max_abs_mel_err_vs_oracle": err}


def measure_c5(dev, hbm_peak, steps=20):
    """BASELINE.json configs[4] ("C5"): LengthRegulator stress, B = 512, Tp = 400, d = 256 fp32, durations U{0..10}
    (expanded length ~2000 frames): scan + scatter kernels against the HBM roofline, output bit-exact vs the oracle."""
    from lightningfastspeech2_b200 import ops
    from oracle import fs2_oracle as O

    g = torch.Generator().manual_seed(5)
    b, tp, d = 512, 400, 256
    x = torch.randn(b, tp, d, generator=g)
    dur = torch.randint(0, 11, (b, tp), generator=g, dtype=torch.int32)
    cap = 2756.25
    xd, dd = x.to(dev), dur.to(dev)
    out, mask = ops.length_regulate(xd, dd, cap)
    l = out.shape[1]
    nref = 48  # bit-exactness against the oracle's torch restatement of model.py:349-370 on the first utterances
    ro, rm = O.length_regulator(x[:nref], dur[:nref].long(), cap)
    w = min(l, ro.shape[1])
    exact = bool(torch.equal(out[:nref, :w].cpu(), ro[:, :w]) and torch.equal(mask[:nref, :w].cpu(), rm[:, :w])
                 and (l >= ro.shape[1]))
    scan = ops.length_regulate_scan(dd, (b, tp))
    for _ in range(3):
        ops.length_regulate_scatter(xd, scan[0], scan[1], l, l)
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    for _ in range(steps):
        scan = ops.length_regulate_scan(dd, (b, tp))
    e[1].record()
    for _ in range(steps):
        ops.length_regulate_scatter(xd, scan[0], scan[1], l, l)
    e[2].record()
    torch.cuda.synchronize()
    ms_scan, ms_scat = e[0].elapsed_time(e[1]) / steps, e[1].elapsed_time(e[2]) / steps
    nbytes = b * tp * (d * 4 + 4) + b * l * (d * 4 + 1)      # x + durations in, frames + mask out
    gbs = nbytes / ((ms_scan + ms_scat) * 1e-3) / 1e9
    return {"workload": f"C5: LengthRegulator B={b}, Tp={tp}, d={d} fp32, durations U{{0..10}} int32 -> {l} frames "
                        "(BASELINE.json configs[4])",
            "ms_scan": ms_scan, "ms_scatter": ms_scat, "algorithmic_bytes": nbytes, "achieved_gbs": gbs,
            "peak_gbs": hbm_peak, "frac_of_hbm_roofline": gbs / hbm_peak,
            "scatter_only_frac": (nbytes - b * tp * 4) / (ms_scat * 1e-3) / 1e9 / hbm_peak,
            "bit_exact_vs_oracle": exact, "checked_utterances": nref,
            "l2": "no flush: the 1.05 GB output exceeds the 126 MB L2"}



def measure_vocoder(dev, mel, tgt_mask, nutt, steps=3, model=None, pinned=None):
    """the HiFi-GAN generator behind the path (reference synthesis/generator.py:160-170 vocodes every
    utterance of the batch, one call each): the first `nutt` utterances of the timed batch's mel output, valid frames
    only, as ONE ragged batch; seeded weights of the reference architecture (the bundled checkpoint cannot travel)."""
    from lightningfastspeech2_b200 import hifigan
    from oracle import hifigan_oracle as HO

    cfg = dict(HO.CONFIG)
    import contextlib

    gen = hifigan.Generator(hifigan.AttrDict(cfg))
    with contextlib.redirect_stdout(sys.stderr):  # (the reference's remove_weight_norm prints; stdout carries ONE JSON line)
        gen.remove_weight_norm()
    sd = synthetic.hifigan_state_dict(cfg, seed=3)
    gen.load_state_dict(sd)
    gen = gen.eval().to(dev)
    lens = (~tgt_mask[:nutt]).sum(1)
    tmax = int(lens.max())
    x = mel[:nutt, :tmax].transpose(1, 2).contiguous()          # (B, 80, T) as the reference feeds it
    with torch.no_grad():
        wav = gen(x, lens)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            wav = gen(x, lens)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    frames = int(lens.sum())
    # the same call in "bf16" mode (single-pass conv1 MMAs; the residual stream stays fp32-exact)
    gen.compute_mode = "bf16"
    with torch.no_grad():
        wav16 = gen(x, lens)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            wav16 = gen(x, lens)
        e1.record()
        torch.cuda.synchronize()
    ms16 = e0.elapsed_time(e1) / steps
    err16 = float((wav16 - wav).abs().max())
    gen.compute_mode = "fp32"
    # the generation loop end to end: phonemes in pinned host memory -> mel -> waveform on the device -> int16 samples of
    # every utterance in pinned host memory (pipeline.SynthesisStream(vocoder=...)), first `nutt` utterances per batch
    stream_res = None
    if model is not None and pinned is not None:
        from lightningfastspeech2_b200.pipeline import SynthesisStream

        sub = {k: v[:nutt].contiguous().pin_memory() for k, v in pinned.items()}
        stream = SynthesisStream(model, vocoder=gen, compact=True)
        stream.collect(stream.submit(sub))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tickets = []
        nsamp = nfr = 0
        for i in range(steps + 1):
            if i < steps:
                tickets.append(stream.submit(sub))
            if i > 0:
                got = stream.collect(tickets[i - 1])
                nsamp += sum(int(w.numel()) for w in got["wav"])
                nfr += sum(got["lengths"])
        dt = time.perf_counter() - t0
        stream_res = {"api": "SynthesisStream(model, vocoder=generator, compact=True).submit / collect", "utterances": nutt,
                      "ms_per_batch": 1e3 * dt / steps, "value": nfr / dt, "unit": UNIT, "samples_per_s": nsamp / dt,
                      "d2h_bytes_per_step": 2 * nsamp // steps + 4 * 80 * nfr // steps}
    # CPU port on the shortest utterance (bounded), and parity on it
    i = int(lens.argmin())
    n = int(lens[i])
    torch.set_num_threads(os.cpu_count())
    t0 = time.perf_counter()
    with torch.no_grad():
        ref = HO.generator(sd, x[i:i + 1, :, :n].cpu())
    cpu_s = time.perf_counter() - t0
    err = float((wav[i, 0, : n * 256].cpu() - ref[0, 0]).abs().max())
    flops_per_frame = 0.0
    ch = cfg["upsample_initial_channel"]
    rate = 1
    flops_per_frame += 2.0 * 80 * ch * 7
    for si, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        c_in, c_out = ch // 2 ** si, ch // 2 ** (si + 1)
        flops_per_frame += 2.0 * c_in * c_out * k * rate          # ConvTranspose1d: k / u taps per output sample
        rate *= u
        flops_per_frame += rate * 2.0 * c_out * c_out * sum(cfg["resblock_kernel_sizes"]) * 6
    flops_per_frame += rate * 2.0 * (ch // 2 ** len(cfg["upsample_rates"])) * 7
    return {"what": "HiFi-GAN v1 generator (reference third_party/hifigan/models.py:112-174; 4 upsamplers x 3 ResBlocks, "
                    "13.9M params, seeded weights) on the mel of the first utterances of the timed batch, ragged batch, "
                    "fp32-parity mode (split-bf16 x3)",
            "utterances": nutt, "mel_frames": frames, "padded_frames": tmax, "ms_per_batch": ms,
            "value": frames / (ms * 1e-3), "unit": UNIT, "samples_per_s": frames * 256 / (ms * 1e-3),
            "realtime_factor": frames * 256 / 22050.0 / (ms * 1e-3),
            "algorithmic_gflop_per_frame": flops_per_frame / 1e9,
            "achieved_tflops": flops_per_frame * frames / (ms * 1e-3) / 1e12,
            "cpu_baseline": {"value": n / cpu_s, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                             "sample": f"the shortest of those utterances ({n} frames), one run of {cpu_s:.1f} s"},
            "max_abs_wav_err_vs_oracle": err,
            "bf16_mode": {"ms_per_batch": ms16, "value": frames / (ms16 * 1e-3), "unit": UNIT,
                          "max_abs_wav_diff_vs_fp32_mode": err16},
            "stream_mel_and_wav": stream_res}


def run_reference(args):
    rank, world, local = dist_env()
    if rank != 0:
        return
    kw = configs.PRESETS[PRESET]
    hp = configs.resolve(kw)
    stats = {v: {"min": -3.0, "max": 3.0, "mean": 0.0, "std": 1.0} for v in hp["variances"]}
    hp["stats"] = stats
    from lightningfastspeech2_b200.fastspeech2.fastspeech2 import FastSpeech2

    model = FastSpeech2(stats=stats, phone2id={f"p{i}": i for i in range(80)}, num_workers=0, **kw)
    sd = synthetic.fill_state_dict(model.state_dict(), seed=0)
    batch = synthetic.make_batch(BATCH, MIN_LEN, MAX_LEN, seed=2)
    # the whole batch the GPU arm times at rank 0 (same config on both arms; ~5.5 s of CPU work per step)
    nutt = BATCH if args.ref_utts is None else min(BATCH, args.ref_utts)
    for _ in range(args.warmup):
        cpu_port_throughput(sd, hp, batch, min(2, nutt))
    times, frames = [], 0
    for _ in range(args.steps):
        _, dt, frames = cpu_port_throughput(sd, hp, batch, nutt)
        times.append(dt)
    total = sum(times)
    value = frames * len(times) / total
    sample = (f"{'all' if nutt == BATCH else 'first ' + str(nutt) + ' of the'} {BATCH} utterances of the C2 batch (seed 2) "
              f"per step = the GPU arm's rank-0 batch; warm-up steps run 2 utterances; oracle/fs2_oracle.py (torch CPU "
              f"conv1d/linear/softmax/layer_norm on {os.cpu_count()} threads), no_grad")
    train = None
    if args.train_cpu_utts > 0:  # the train-step leg of the metric on the CPU port, same global batch, bounded sample
        _, sd4, hp4 = build_train_model(None)
        tb = train_batch(hp4, 0, 1)
        secs, cfr = cpu_port_train_step(sd4, hp4, tb, args.train_cpu_utts)
        frames4 = int(tb["duration"].sum())
        train = {"metric": "ms/step (train)", "ms_per_step": secs * 1e3 * frames4 / max(cfr, 1), "impl": "reference",
                 "cores": os.cpu_count(), "kind": "port", "valid_frames_per_step": frames4,
                 "sample": f"first {args.train_cpu_utts} utterances of the C4 batch ({cfr} frames): one oracle train step "
                           f"(forward + loss + autograd backward + AdamW/Noam) of {secs:.1f} s, scaled to {frames4} frames",
                 "config": {"workload": TRAIN_WORKLOAD}}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "preset": PRESET, "utterances_per_gpu": BATCH, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if train is not None:
        line["train"] = train
    print(json.dumps(line))


def run_lfs2(args):
    import torch.distributed as dist

    from lightningfastspeech2_b200 import _lib, ops

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the lfs2 path has no CPU fallback); "
                         "use --impl reference for the CPU port")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    model, sd, hp = build_model(dev)
    host_batch = synthetic.make_batch(BATCH, MIN_LEN, MAX_LEN, seed=2 + rank)
    pinned = {k: v.pin_memory() for k, v in host_batch.items() if k in ("phones", "speaker")}
    resident = {k: v.to(dev) for k, v in pinned.items()}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        with torch.no_grad():
            return model(resident, inference=True)

    frames = None
    for _ in range(max(args.warmup, 3)):
        r = step_resident()
    frames = int((~r["tgt_mask"]).sum())
    mel_shape = tuple(r["mel"].shape)
    tp = resident["phones"].shape[1]

    # ---- timed region: inputs resident in HBM -------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    calls0 = _lib.CALLS
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.CALLS - calls0

    # ---- e2e: host inputs, pinned H2D + D2H of mel / mask inside the timed region ---------
    mel_host = torch.empty(mel_shape, dtype=torch.float32).pin_memory()
    mask_host = torch.empty(mel_shape[:2], dtype=torch.bool).pin_memory()

    def step_e2e():
        with torch.no_grad():
            out = model(pinned, inference=True)
        mel_host.copy_(out["mel"], non_blocking=True)
        mask_host.copy_(out["tgt_mask"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return int((~mask_host).sum())

    step_e2e()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_frames = 0
    for _ in range(args.steps):
        e2e_frames += step_e2e()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    # ---- e2e, pipelined: the same per-step work (pinned H2D of the inputs, D2H of mel + mask into pinned memory,
    #      host reads the mask), with the read-back of step i on a second stream overlapping step i+1's kernels
    #      (lightningfastspeech2_b200.pipeline.SynthesisStream, depth 2) ----------------------------------------
    from lightningfastspeech2_b200.pipeline import SynthesisStream

    # (a) padded read-back: the whole (B, L, 80) mel + mask, as model(batch) returns them
    pipe = SynthesisStream(model, depth=2)
    for _ in range(4):  # warm-up: both slots allocate their pinned buffers, the allocator reaches its steady state
        pipe.collect(pipe.submit(pinned))
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    padded_frames, prev = 0, None
    for _ in range(args.steps):
        tk = pipe.submit(pinned)
        if prev is not None:
            padded_frames += int((~pipe.collect(prev)["tgt_mask"]).sum())
        prev = tk
    padded_frames += int((~pipe.collect(prev)["tgt_mask"]).sum())
    p1.record()
    barrier()
    ms_piped_padded = p0.elapsed_time(p1)
    # (b) compact read-back (the headline e2e): every utterance's mel cut at its own length -- what the reference's
    #     caller keeps (synthesis/generator.py:164-170) -- packed on the device, one transfer of sum(frames) rows
    pipe = SynthesisStream(model, depth=2, compact=True)
    for _ in range(4):
        pipe.collect(pipe.submit(pinned))
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    piped_frames, prev, d2h_compact = 0, None, 0
    for _ in range(args.steps):
        tk = pipe.submit(pinned)
        if prev is not None:
            got = pipe.collect(prev)
            piped_frames += sum(got["lengths"])
            d2h_compact = sum(m.numel() for m in got["mel"]) * 4 + 8 * len(got["lengths"])
        prev = tk
    got = pipe.collect(prev)
    piped_frames += sum(got["lengths"])
    d2h_compact = sum(m.numel() for m in got["mel"]) * 4 + 8 * len(got["lengths"])
    assert piped_frames == padded_frames, (piped_frames, padded_frames)
    p1.record()
    barrier()
    ms_piped = p0.elapsed_time(p1)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel CUDA-event pass (same inputs, after the timed region) -----------------
    ops.PROFILE = {}
    for _ in range(min(args.steps, 3)):
        step_resident()
    torch.cuda.synchronize()
    prof = ops.collect_profile()
    ops.PROFILE = None

    # ---- reduce over ranks ------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_piped, ms_piped_padded], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_piped, ms_piped_padded = float(t[0]), float(t[1]), float(t[2]), float(t[3])
        c = torch.tensor([frames, e2e_frames, piped_frames], device=dev, dtype=torch.int64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        frames_all, e2e_all, piped_all = int(c[0]), int(c[1]), int(c[2])
    else:
        frames_all, e2e_all, piped_all = frames, e2e_frames, piped_frames

    # The sections below are SECONDARY numbers.  Each runs under try/except with local synchronisation only, and its
    # cross-rank reduction happens afterwards through reduce_section() on every rank, so that a failure in one of them
    # can neither take the headline line down nor leave another rank waiting in a collective.
    def local_sync():
        torch.cuda.synchronize()

    def reduce_section(ok, times, counts):
        """-> (every rank succeeded, max over ranks of the times, sum over ranks of the counts)"""
        if world == 1:
            return ok, times, counts
        tmax = torch.tensor(list(times) or [0.0], device=dev, dtype=torch.float64)
        tsum = torch.tensor([1.0 if ok else 0.0] + list(counts), device=dev, dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        return bool(tsum[0] == world), [float(x) for x in tmax[:len(times)]], [int(x) for x in tsum[1:]]

    def timed(fn, steps):
        """K calls of fn between two events on the current stream -> (ms, last result)"""
        local_sync()
        e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_a.record()
        out = None
        for _ in range(steps):
            out = fn()
        e_b.record()
        local_sync()
        return e_a.elapsed_time(e_b), out

    errors = {}

    # ---- bf16 mode (single-pass bf16 MMA operands, fp32 accumulate / residual / LayerNorm / softmax; tolerance
    #      1e-2 per BASELINE.json) on the same batch: secondary number, same timing protocol -------------------
    ok_bf16, ms_bf16, frames_bf16, prof_bf16 = True, 0.0, 0, {}
    try:
        model.set_compute_mode("bf16")
        for _ in range(3):
            step_resident()
        ms_bf16, rb = timed(step_resident, args.steps)
        frames_bf16 = int((~rb["tgt_mask"]).sum())
        ops.PROFILE = {}
        step_resident()
        prof_bf16 = ops.collect_profile()
    except Exception as exc:  # noqa: BLE001
        ok_bf16, errors["bf16_mode"] = False, repr(exc)[:300]
    finally:
        ops.PROFILE = None
        model.set_compute_mode("fp32")
    ok_bf16, (ms_bf16,), (frames_bf16,) = reduce_section(ok_bf16, [ms_bf16], [frames_bf16])

    # ---- length-bucketed synthesis of the SAME batch (valid frames bit-identical, PAD frames zero) ----------
    bucketed = []
    for nb in args.buckets:
        ok_b, ms_b, ms_be, fr_b, launches_b = True, 0.0, 0.0, 0, 0
        try:
            model.length_buckets = nb
            for _ in range(3):
                step_resident()
            calls_b = _lib.CALLS
            ms_b, _ = timed(step_resident, args.steps)
            launches_b = _lib.CALLS - calls_b
            step_e2e()
            local_sync()
            e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e_a.record()
            for _ in range(args.steps):
                fr_b += step_e2e()
            e_b.record()
            local_sync()
            ms_be = e_a.elapsed_time(e_b)
        except Exception as exc:  # noqa: BLE001
            ok_b, errors[f"bucketed_{nb}"] = False, repr(exc)[:300]
        finally:
            model.length_buckets = 1
        ok_b, (ms_b, ms_be), (fr_b,) = reduce_section(ok_b, [ms_b, ms_be], [fr_b])
        if ok_b:
            bucketed.append({"length_buckets": nb, "ms_per_step": ms_b / args.steps, "gpu_launches": launches_b,
                             "frames_all_ranks_e2e": fr_b, "ms_per_step_e2e": ms_be / args.steps})

    # ---- PAD-row skipping synthesis of the SAME batch (model.skip_pad_rows: kernels run only over the 128-row tiles
    #      before an utterance's end + conv halo; valid frames bit-identical, masked frames zero) ------------------
    pad_skip = None
    ok_s, ms_s, ms_sp, ms_sb, fr_s, fr_sp, launches_s, same_s, prof_s, rows_s = True, 0.0, 0.0, 0.0, 0, 0, 0, False, {}, 1.0
    try:
        with torch.no_grad():
            full_out = model(resident, inference=True)
        model.skip_pad_rows = True
        rs = step_resident()
        valid_s = ~full_out["tgt_mask"]
        fr_s = int(valid_s.sum())
        same_s = bool(torch.equal(rs["tgt_mask"], full_out["tgt_mask"]) and
                      torch.equal(rs["mel"][valid_s], full_out["mel"][valid_s]))
        kept = torch.clamp((valid_s.sum(1) + sum(l.halo() for l in model.decoder.layers) + 127) // 128 * 128,
                           max=valid_s.shape[1]).sum().item()
        rows_s = kept / float(valid_s.numel())
        del full_out, rs
        for _ in range(8):  # the allocator needs a few steps to settle on the new (smaller) working set
            step_resident()
        calls_s = _lib.CALLS
        ms_s, _ = timed(step_resident, args.steps)
        launches_s = (_lib.CALLS - calls_s) // max(args.steps, 1)
        for _ in range(4):
            pipe.collect(pipe.submit(pinned))
        local_sync()
        e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_a.record()
        prev = None
        for _ in range(args.steps):
            tk = pipe.submit(pinned)
            if prev is not None:
                fr_sp += sum(pipe.collect(prev)["lengths"])
            prev = tk
        fr_sp += sum(pipe.collect(prev)["lengths"])
        e_b.record()
        local_sync()
        ms_sp = e_a.elapsed_time(e_b)
        ops.PROFILE = {}
        step_resident()
        prof_s = ops.collect_profile()
        ops.PROFILE = None
        model.set_compute_mode("bf16")
        for _ in range(3):
            step_resident()
        ms_sb, _ = timed(step_resident, args.steps)
    except Exception as exc:  # noqa: BLE001
        ok_s, errors["pad_skip"] = False, repr(exc)[:300]
    finally:
        ops.PROFILE = None
        model.skip_pad_rows = False
        model.set_compute_mode("fp32")
    ok_s, (ms_s, ms_sp, ms_sb), (fr_s, fr_sp, n_same) = reduce_section(ok_s, [ms_s, ms_sp, ms_sb], [fr_s, fr_sp, int(same_s)])
    if ok_s:
        pad_skip = {"ms": ms_s, "ms_piped": ms_sp, "ms_bf16": ms_sb, "frames": fr_s, "frames_piped": fr_sp,
                    "launches": launches_s, "identical": n_same == world, "prof": prof_s, "rows": rows_s}

    # ---- parity of the timed batch against the oracle (rank 0; the first utterances, padded like the timed tensors) ----
    parity = {}
    if rank == 0 and args.parity_utts > 0:
        try:
            parity["c2"] = parity_vs_oracle(model, sd, hp, host_batch, args.parity_utts,
                                            {"fp32": ("fp32", False), "bf16": ("bf16", False),
                                             "fp32_skip_pad_rows": ("fp32", True)})
        except Exception as exc:  # noqa: BLE001
            errors["parity_c2"] = repr(exc)[:300]

    # ---- N1: the vocoder behind the path, on this batch's mel (rank 0) ----
    vocoder = None
    if rank == 0 and args.vocoder_utts > 0:
        try:
            with torch.no_grad():
                rv = model(resident, inference=True)
            vocoder = measure_vocoder(dev, rv["mel"], rv["tgt_mask"], args.vocoder_utts, model=model, pinned=pinned)
            del rv
        except Exception as exc:  # noqa: BLE001
            errors["vocoder"] = repr(exc)[:300]
        torch.cuda.empty_cache()

    # ---- BASELINE.json configs[2] ("C3"): 76 M-parameter model, bf16 synthesis, 32 utterances per GPU ----------
    c3 = None
    if args.c3_steps > 0:
        model = None
        torch.cuda.empty_cache()
        ok3, ms3, fr3, shape3, buck3, skip3 = True, 0.0, 0, [], [], None
        try:
            m3, sd3, hp3 = build_model(dev, preset="C3")
            m3.set_compute_mode("bf16")
            hb3 = synthetic.make_batch(32, MIN_LEN, MAX_LEN, seed=200 + rank)
            b3 = {k: v.to(dev) for k, v in hb3.items() if k in ("phones", "speaker")}

            def step3():
                with torch.no_grad():
                    return m3(b3, inference=True)

            for _ in range(3):
                step3()
            ms3, r3 = timed(step3, args.c3_steps)
            ms3 /= args.c3_steps
            fr3 = int((~r3["tgt_mask"]).sum())
            shape3 = list(r3["mel"].shape)
            # the same call as length-sorted sub-batches (model.length_buckets: valid frames bit-identical)
            for nb3 in (2, 3, 4):
                m3.length_buckets = nb3
                for _ in range(3):
                    step3()
                t3, rb3 = timed(step3, args.c3_steps)
                same = bool(torch.equal(rb3["mel"][~r3["tgt_mask"]], r3["mel"][~r3["tgt_mask"]]))
                buck3.append({"length_buckets": nb3, "ms_per_step": t3 / args.c3_steps,
                              "value": fr3 / (t3 / args.c3_steps * 1e-3), "unit": UNIT + " (rank 0's shard)",
                              "valid_frames_bit_identical": same})
            m3.length_buckets = 1
            # and with PAD-row skipping in one launch per kernel (model.skip_pad_rows, the wide row-limited block)
            m3.skip_pad_rows = True
            for _ in range(3):
                step3()
            t3, rb3 = timed(step3, args.c3_steps)
            m3.skip_pad_rows = False
            skip3 = {"ms_per_step": t3 / args.c3_steps, "value": fr3 / (t3 / args.c3_steps * 1e-3),
                     "unit": UNIT + " (rank 0's shard)",
                     "valid_frames_bit_identical": bool(torch.equal(rb3["mel"][~r3["tgt_mask"]], r3["mel"][~r3["tgt_mask"]]))}
            if rank == 0 and args.parity_utts > 0:
                try:
                    parity["c3"] = parity_vs_oracle(m3, sd3, hp3, hb3, max(1, args.parity_utts // 4),
                                                    {"bf16": ("bf16", False), "fp32": ("fp32", False)})
                except Exception as exc:  # noqa: BLE001
                    errors["parity_c3"] = repr(exc)[:300]
            del m3, r3
        except Exception as exc:  # noqa: BLE001
            ok3, errors["c3_bf16"] = False, repr(exc)[:300]
        torch.cuda.empty_cache()
        ok3, (ms3,), (fr3,) = reduce_section(ok3, [ms3], [fr3])
        if ok3:
            c3 = {"workload": "C3: 76M-parameter model (d=768, head_dim 384, 4 enc + 5 dec FFTBlocks, 3 variances), bf16 mode, "
                              "32 utterances per GPU, phoneme len U[32,512] (BASELINE.json configs[2]: 256 utterances over 8 GPUs)",
                  "value": fr3 / (ms3 * 1e-3), "unit": UNIT, "ms_per_step": ms3, "valid_frames_per_step": fr3,
                  "mel_shape_rank0": shape3, "steps": args.c3_steps}
            if buck3:  # rank 0's own timing of the bucketed variants (no cross-rank reduction)
                c3["bucketed"] = buck3
            if skip3:
                c3["pad_skip"] = skip3

    train = None
    if args.train_steps > 0:
        model = None
        torch.cuda.empty_cache()
        try:
            train = measure_train(args, dev, rank, world, barrier)
        except Exception as exc:  # noqa: BLE001
            errors["train"] = repr(exc)[:300]

    c1 = c5 = None
    if rank == 0 and args.c1_steps > 0:
        try:
            c1 = measure_c1(dev, args.c1_steps)
        except Exception as exc:  # noqa: BLE001
            errors["c1"] = repr(exc)[:300]
    if rank == 0 and args.c5_steps > 0:
        try:
            c5 = measure_c5(dev, peaks()["hbm"], args.c5_steps)
        except Exception as exc:  # noqa: BLE001
            errors["c5"] = repr(exc)[:300]
        torch.cuda.empty_cache()

    if rank == 0:
        pk = peaks()
        total_ms = sum(v["ms"] for v in prof.values())
        top_name = max(prof, key=lambda k: prof[k]["ms"])
        top = prof[top_name]
        per_launch_s = top["ms"] * 1e-3 / top["launches"]
        if top["bound"] == "tensor":
            achieved = top["flops"] / top["launches"] / per_launch_s / 1e12
            # the kernel is timed inside a long step (back-to-back launches, power-capped clocks): the SUSTAINED
            # dense bf16 figure of MEASURED_PEAKS.json is the denominator, as the profiling recipe prescribes
            peak, unit = pk["tensor_sustained"], "TFLOP/s"
        else:
            achieved = top["bytes"] / top["launches"] / per_launch_s / 1e9
            peak, unit = pk["hbm"], "GB/s"
        roofline = {"kernel": top_name, "bound": "hbm" if top["bound"] == "hbm" else "tensor",
                    "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                    # fp32-parity mode issues 3 bf16 MMA passes (hi.hi + lo.hi + hi.lo) per algorithmic product:
                    # `achieved`/`frac` count each product ONCE; the tensor pipe executes 3x that
                    "mma_passes": round(top["issued_flops"] / top["flops"]) if top["bound"] == "tensor" and top["flops"] else None,
                    "issued_frac": (top["issued_flops"] / top["launches"] / per_launch_s / 1e12 / peak)
                    if top["bound"] == "tensor" else None,
                    "traffic": ncu_traffic(top_name),
                    "peak_source": pk["src"] + (" (bf16_tflops_sustained)" if top["bound"] == "tensor" else " (hbm_gbs)"), "us_per_launch": per_launch_s * 1e6,
                    "share_of_step": top["ms"] / total_ms,
                    "kernel_shares": {k: round(v["ms"] / total_ms, 4) for k, v in sorted(
                        prof.items(), key=lambda kv: -kv[1]["ms"])},
                    # every kernel of the step against ITS roofline (algorithmic bytes / flops of SURVEY 8d, CUDA
                    # events on the launching stream): frac = max(bytes/t / HBM peak, flops/t / tensor peak)
                    # frac = algorithmic (each product counted once); frac_issued = against the roofline of the mode the
                    # kernel runs in: a split-bf16 product is 3 MMA passes, so its tensor roofline is a third of the peak
                    "per_kernel": {k: {"launches": v["launches"], "ms": round(v["ms"], 4), "bound": v["bound"],
                                       "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["ms"] > 0 else 0.0,
                                       "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["ms"] > 0 else 0.0,
                                       "frac": round(max(v["bytes"] / (v["ms"] * 1e-3) / 1e9 / pk["hbm"],
                                                         v["flops"] / (v["ms"] * 1e-3) / 1e12 / pk["tensor_sustained"]), 3)
                                       if v["ms"] > 0 else 0.0,
                                       "frac_issued": round(max(v["bytes"] / (v["ms"] * 1e-3) / 1e9 / pk["hbm"],
                                                                v["issued_flops"] / (v["ms"] * 1e-3) / 1e12 / pk["tensor_sustained"]), 3)
                                       if v["ms"] > 0 else 0.0}
                                   for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:12]}}
        # bounded CPU sample of the same workload: grow the sub-batch until one run takes >= ~10 s of CPU work
        nutt = max(1, args.ref_utts or 16)
        while True:
            cpu_fps, cpu_s, cpu_frames = cpu_port_throughput(sd, hp, host_batch, nutt)
            if cpu_s >= 10.0 or nutt >= min(BATCH, args.ref_utts_max):
                break
            nutt = min(BATCH, args.ref_utts_max, nutt * (4 if cpu_s < 1.5 else 2))
        sample = (f"first {nutt} of the {BATCH} utterances of the same batch (padded to their own max length), "
                  f"1 run of {cpu_s:.1f} s ({cpu_frames} valid frames), torch CPU ops on {os.cpu_count()} threads")
        value = frames_all * args.steps / (ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "preset": PRESET, "utterances_per_gpu": BATCH, "padded_phones": tp,
                       "mel_shape_rank0": list(mel_shape), "valid_frames_per_step": frames_all,
                       "l2": "no flush: every activation tensor of a step (>=168 MB) exceeds the 126 MB L2",
                       "parallelism": f"utterance-sharded x{world}, no collective"},
            "clocks": clocks,
            # end to end through the public API with HOST buffers: every step copies its inputs from pinned host memory
            # and its mel + mask back to pinned host memory, and the host reads the mask.  `value` = the generation-loop
            # API (pipeline.SynthesisStream, depth 2: step i's read-back overlaps step i+1's kernels) reading back every
            # utterance's mel cut at its own length (compact=True: what the reference's caller keeps, generator.py:164-170);
            # `padded` = the same stream reading back the whole padded mel + mask; `sequential` = a blocking read-back of
            # the padded mel after every model(batch) call.
            "e2e": {"value": piped_all / (ms_piped * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in pinned.values()) * world,
                    "d2h_bytes_per_step": (d2h_compact + 8 * BATCH) * world,
                    "ms_per_step": ms_piped / args.steps,
                    "api": "lightningfastspeech2_b200.pipeline.SynthesisStream(model, compact=True).submit / collect",
                    "padded": {"value": piped_all / (ms_piped_padded * 1e-3), "ms_per_step": ms_piped_padded / args.steps,
                               "d2h_bytes_per_step": (mel_host.numel() * 4 + mask_host.numel() + 8) * world,
                               "api": "SynthesisStream(model).submit / collect"},
                    "sequential": {"value": e2e_all / (ms_e2e * 1e-3), "ms_per_step": ms_e2e / args.steps,
                                   "api": "model(batch, inference=True); mel.cpu()"}},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_fps, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                             "sample": sample},
        }
        if errors:
            line["errors"] = errors
        # the short sections first: the driver keeps only the tail of a long line
        if parity:
            line["parity_check"] = parity
        if c3 is not None:
            line["c3_bf16"] = c3
        if train is not None:
            line["train"] = train
        if c1 is not None:
            line["c1"] = c1
        if c5 is not None:
            line["c5_length_regulator"] = c5
        if vocoder is not None:
            line["vocoder_hifigan"] = vocoder
        tot_b = sum(v["ms"] for v in prof_bf16.values()) or 1.0
        top_b = max(prof_bf16, key=lambda k: prof_bf16[k]["ms"]) if prof_bf16 else None
        tb = prof_bf16[top_b] if top_b else None
        if ok_bf16 and tb:
            line["bf16_mode"] = {
                "what": "same batch and API call with model.set_compute_mode('bf16'): single-pass bf16 MMA operands, fp32 "
                        "accumulation / residual / LayerNorm / softmax (mel within 1e-2 of the fp32 reference)",
                "value": frames_bf16 * args.steps / (ms_bf16 * 1e-3), "unit": UNIT, "ms_per_step": ms_bf16 / args.steps,
                "dominant_kernel": top_b, "share_of_step": tb["ms"] / tot_b,
                "achieved_tflops": tb["flops"] / (tb["ms"] * 1e-3) / 1e12,
                "achieved_gbs": tb["bytes"] / (tb["ms"] * 1e-3) / 1e9,
                "kernel_shares": {k: round(v["ms"] / tot_b, 4)
                                  for k, v in sorted(prof_bf16.items(), key=lambda kv: -kv[1]["ms"])[:8]}}
        if bucketed:
            line["bucketed"] = {
                "what": "same batch, same API call with model.length_buckets = n: length-sorted sub-batches padded to "
                        "their own longest utterance + the conv halo; valid frames are bit-identical to the full padded "
                        "batch (tests/test_gpu_forward.py), frames masked by tgt_mask come back as zeros",
                "runs": [{"length_buckets": b["length_buckets"], "ms_per_step": b["ms_per_step"],
                          "value": frames_all * args.steps / (b["ms_per_step"] * args.steps * 1e-3),
                          "e2e_value": b["frames_all_ranks_e2e"] / (b["ms_per_step_e2e"] * args.steps * 1e-3),
                          "gpu_launches": b["gpu_launches"], "unit": UNIT} for b in bucketed]}
        if pad_skip is not None:
            ps = pad_skip
            tot_s = sum(v["ms"] for v in ps["prof"].values()) or 1.0
            line["pad_skip"] = {
                "what": "same batch, same API call with model.skip_pad_rows = True: every encoder/decoder kernel runs only "
                        "over the 128-row tiles that start before an utterance's end + the downstream conv half-widths "
                        "(PAD rows are never attention keys, so rows farther out cannot reach a valid frame); the headline "
                        "`value` above does NOT use it and computes every PAD row like the reference",
                "valid_frames_bit_identical_to_headline_path": ps["identical"],
                "decoder_rows_computed_frac": round(ps["rows"], 4),
                "value": ps["frames"] * args.steps / (ps["ms"] * 1e-3), "unit": UNIT, "ms_per_step": ps["ms"] / args.steps,
                "e2e_value": ps["frames_piped"] / (ps["ms_piped"] * 1e-3), "e2e_ms_per_step": ps["ms_piped"] / args.steps,
                "e2e_api": "SynthesisStream(model, compact=True).submit / collect (host inputs, every utterance's valid frames read back to pinned memory)",
                "bf16_mode_value": ps["frames"] * args.steps / (ps["ms_bf16"] * 1e-3) if ps["ms_bf16"] else None,
                "gpu_launches_per_step": ps["launches"],
                "kernel_ms": {k: round(v["ms"], 4) for k, v in sorted(ps["prof"].items(), key=lambda kv: -kv[1]["ms"])[:8]},
                "kernel_shares": {k: round(v["ms"] / tot_s, 4)
                                  for k, v in sorted(ps["prof"].items(), key=lambda kv: -kv[1]["ms"])[:8]}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="lfs2", choices=["lfs2", "reference"])
    ap.add_argument("--ref-utts", type=int, default=None,
                    help="utterances in the CPU sample: --impl reference times all 64 by default; the GPU arm's cpu_baseline "
                         "starts at 16 and grows until one run takes ~10 s")
    ap.add_argument("--ref-utts-max", type=int, default=64, help="upper bound of the adaptive CPU sample")
    ap.add_argument("--buckets", type=int, nargs="*", default=[2, 4], help="length_buckets values of the 'bucketed' runs")
    ap.add_argument("--c3-steps", type=int, default=10, help="timed C3 (76M, bf16) synthesis steps under 'c3_bf16' (0 = skip)")
    ap.add_argument("--parity-utts", type=int, default=16,
                    help="utterances of the timed C2 batch compared with the oracle under 'parity_check' (C3: a quarter; 0 = skip)")
    ap.add_argument("--c1-steps", type=int, default=30, help="timed C1 (1 x 128 phonemes, dense k=9) calls under 'c1' (0 = skip)")
    ap.add_argument("--c5-steps", type=int, default=20, help="timed C5 LengthRegulator launches under 'c5_length_regulator' (0 = skip)")
    ap.add_argument("--vocoder-utts", type=int, default=8,
                    help="utterances of the timed batch vocoded by the HiFi-GAN generator under 'vocoder_hifigan' (0 = skip)")
    ap.add_argument("--train-steps", type=int, default=5, help="timed C4 train steps reported under 'train' (0 = skip)")
    ap.add_argument("--train-mode", default="fp32", choices=["simt", "fp32", "bf16"])
    ap.add_argument("--train-buckets", type=int, nargs="*", default=[2, 3, 4],
                    help="model.train_length_buckets values timed under train.length_buckets")
    ap.add_argument("--train-cpu-utts", type=int, default=2, help="utterances in the CPU train-step sample (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_lfs2(args)


if __name__ == "__main__":
    main()
