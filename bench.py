#!/usr/bin/env python
"""bench.py - STE-GAN hot-path benchmark (BASELINE.json: GAN train samples/s, configs[1]).

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

One "step" = one full GAN train step (train.py:165-268 without the encoder losses) on a
synthetic batch of 16 samples per GPU, 100 unit frames -> 1600x8 EMG samples, bf16 tensor-core
mode, weak scaling (per-GPU batch fixed).  Prints ONE JSON line (rank 0).

  value   whole-job samples/s with inputs resident in HBM, CUDA-graph replay, device-timed
  e2e     the same through GanTrainer.step_graph() with pinned HOST inputs (H2D inside the timed
          region) and a host read of the loss slots every step
  roofline  dominant kernel (tcgen05 implicit-GEMM conv) timed live with CUDA events in an
          instrumented eager step: algorithmic FLOPs / event time vs MEASURED_PEAKS.json
  cpu_baseline  the oracle port of the reference step on the host cores (bounded sample)
  --impl reference  times that CPU path alone (rank 0), same metric / unit / config
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH_PER_GPU, FRAMES, UNIT_DIM, HOP, CHANNELS = 16, 100, 256, 16, 8
STEP_GFLOP_PER_SAMPLE = 94.7          # G 41.42 + D fwd 4x5.919 + D bwd 2x11.84 + D dgrad 5.92
INFER_FRAMES = 1500                   # configs[3]: 30 s utterances
G_FWD_GFLOP_PER_FRAME = 13.807 / 100


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(tflops=float(d["bf16_tflops_sustained"]), burst=float(d["bf16_tflops"]), hbm=float(d["hbm_gbs"]),
                    source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(tflops=1400.0, burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill(); out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def config_dict(world: int):
    return {"workload": "STE-GAN base train step (G + DiscriminatorSmall fwd/bwd, LSGAN + 15*multi-TD + 7*FM, 2x AdamW), "
                        "bf16, batch 16/GPU, 100 unit frames -> 1600x8 EMG samples (BASELINE.json configs[1])",
            "batch_per_gpu": BATCH_PER_GPU, "global_batch": BATCH_PER_GPU * world, "unit_frames": FRAMES,
            "emg_samples": FRAMES * HOP, "parallelism": f"dp{world}",
            "l2": "per-step working set (weights 212 MB + activations) exceeds the 126 MB L2; inputs rotate over 4 batches"}


# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int, budget_s: float = 200.0):
    """The reference's CPU path (oracle port of train.py:165-268, fp32, all host threads)."""
    import torch
    from oracle import ste_gan_oracle as O
    from ste_gan_b200.models.discriminator import DiscriminatorSmall
    from ste_gan_b200.models.generator import EMGGeneratorGanTTS
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0); g = EMGGeneratorGanTTS("SPEECH_UNITS", 256, 17, 8)
    torch.manual_seed(0); d = DiscriminatorSmall(8)
    ot = O.OracleTrainer({k: v for k, v in g.state_dict().items()}, {k: v for k, v in d.state_dict().items()}, small=True)
    batch = BATCH_PER_GPU
    t0 = time.perf_counter()
    ot.step(*O.synthetic_batch(batch, FRAMES, seed=0))
    first = time.perf_counter() - t0
    if first * (steps + max(0, warmup - 1)) > budget_s:      # bound the sample: fewer samples per step
        batch = max(1, int(batch * budget_s / (first * (steps + max(0, warmup - 1)))))
    for i in range(max(0, warmup - 1)):
        ot.step(*O.synthetic_batch(batch, FRAMES, seed=1 + i))
    times = []
    for i in range(steps):
        b = O.synthetic_batch(batch, FRAMES, seed=100 + i)
        t0 = time.perf_counter(); ot.step(*b); times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    return dict(value=batch / (ms / 1e3), ms_per_step=ms, batch=batch, cores=torch.get_num_threads(),
                sample=f"{steps} timed steps of batch {batch} x {FRAMES} frames (fp32, torch CPU ops, {torch.get_num_threads()} threads)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup)
    line = {"impl": "reference", "metric": "GAN train samples/s", "value": r["value"], "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(1),
            "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from oracle import ste_gan_oracle as O          # synthetic batch generator (shared with the tests)
    from ste_gan_b200 import _lib, ops
    from ste_gan_b200.dist import init_from_env
    from ste_gan_b200.inference import UtteranceGenerator
    from ste_gan_b200.models.discriminator import DiscriminatorSmall
    from ste_gan_b200.models.generator import EMGGeneratorGanTTS
    from ste_gan_b200.trainer import GanTrainer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback (use --impl reference for the CPU arm)")
    rank, world, local = init_from_env("nccl")
    if world != args.gpus:
        raise SystemExit(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; launch N>1 with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load(build_if_missing=False)
    peaks = load_peaks()

    torch.manual_seed(0); g = EMGGeneratorGanTTS("SPEECH_UNITS", 256, 17, 8).to(dev)
    torch.manual_seed(0); d = DiscriminatorSmall(8).to(dev)
    tr = GanTrainer(g, d, precision="bf16")
    tr.reducer.broadcast(tr.G.flat); tr.reducer.broadcast(tr.D.flat)

    nb = 4
    host = [O.synthetic_batch(BATCH_PER_GPU, FRAMES, seed=1000 * rank + i) for i in range(nb)]
    pinned = [tuple(t.pin_memory() for t in b) for b in host]
    devb = [tuple(t.to(dev) for t in b) for b in host]

    c0 = lib.stg_launch_count()
    tr.capture(BATCH_PER_GPU, FRAMES, UNIT_DIM, HOP, CHANNELS)
    # capture() = 2 eager warm-up steps + one capture pass: launches per step = a third of the delta
    launches_per_step = (lib.stg_launch_count() - c0) // 3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for i in range(max(3, args.warmup)):
        tr.step_graph(*devb[i % nb])
    sampler = ClockSampler(local)
    # ---- timed region 1: device-resident inputs
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        tr.step_graph(*devb[i % nb])
    tr.flush()                  # the last step's deferred all-reduce + generator optimiser belong to the timed region
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    ms_step = ms_total / args.steps
    value = BATCH_PER_GPU * world / (ms_step / 1e3)
    final_losses = tr.losses()

    # ---- timed region 2: end to end (pinned host inputs, loss read back every step)
    for i in range(2):
        tr.step_graph(*pinned[i % nb]); tr.losses()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        tr.step_graph(*pinned[i % nb])
        tr.slots.tolist()
    tr.flush()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    h2d = sum(t.numel() * t.element_size() for t in pinned[0])
    e2e = {"value": BATCH_PER_GPU * world / (ms_e2e / 1e3), "unit": "samples/s", "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": tr.slots.numel() * 4}

    # ---- roofline: instrumented eager step (every conv / wgrad launch bracketed by CUDA events)
    # (every rank runs these steps - they contain the gradient all-reduce - but only rank 0 keeps the numbers)
    roofline, by_kernel = None, []
    # A 60 ms device-side sleep ahead of each instrumented step lets the host enqueue the whole step before the GPU
    # starts it, so the event pairs time back-to-back kernels rather than host launch gaps.
    # The instrumented steps run on ONE stream (concurrent_d = False): each kernel is timed alone on the chip, which is
    # what a per-kernel roofline fraction means.  In the timed (graph) step the same launches are spread over up to 12
    # streams and time-share the SMs, so their per-launch durations there are longer than the kernel's own.
    tr.concurrent_d = False
    tr.step(*devb[0])
    ops.profile = []
    for j in (1, 2):
        torch.cuda._sleep(int(0.06 * 1.9e9))
        tr.step(*devb[j])
    torch.cuda.synchronize()
    tr.concurrent_d = True
    prof, ops.profile = ops.profile, None
    if rank == 0:
        groups = {}
        for p in prof:
            key = {"tcgen05": ("conv_tc_kernel", "wgrad_tc_kernel"), "simt": ("conv_simt_kernel", "wgrad_simt_kernel"),
                   "matvec": ("c1_conv_kernels", "c1_wgrad_kernel")}[p["engine"]][1 if p["kind"] == "wgrad" else 0]
            gk = groups.setdefault(key, dict(kernel=key, launches=0, ms=0.0, flops=0.0, bytes=0.0))
            gk["launches"] += 1; gk["ms"] += p["events"][0].elapsed_time(p["events"][1]); gk["flops"] += p["flops"]
            gk["bytes"] += p["bytes"]
        n_steps_prof = 2
        for gk in groups.values():
            gk["tflops"] = gk["flops"] / (gk["ms"] * 1e-3) / 1e12
            gk["launches_per_step"] = gk["launches"] // n_steps_prof
            gk["ms_per_step"] = gk["ms"] / n_steps_prof
            gk["gflop_per_step"] = gk["flops"] / n_steps_prof / 1e9
            by_kernel.append({k: (round(v, 4) if isinstance(v, float) else v) for k, v in gk.items()
                              if k in ("kernel", "launches_per_step", "ms_per_step", "gflop_per_step", "tflops")})
        top = max(groups.values(), key=lambda x: x["ms"])
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")   # written from the ncu --set full capture
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(top["kernel"])
        roofline = {"bound": "tensor", "kernel": top["kernel"], "achieved": round(top["tflops"], 2),
                    "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": round(top["tflops"] / peaks["tflops"], 4),
                    "traffic": traffic, "peak_source": peaks["source"],
                    "timing": "CUDA events around every launch of two eager steps on one stream (kernel alone on the chip)",
                    "launches_per_step": top["launches"] // n_steps_prof,
                    "avg_launch_us": round(1e3 * top["ms"] / top["launches"], 2),
                    "algorithmic_gflop_per_launch": round(top["flops"] / top["launches"] / 1e9, 3),
                    "step_tflops": round(STEP_GFLOP_PER_SAMPLE * BATCH_PER_GPU / ms_step, 2),
                    "step_frac": round(STEP_GFLOP_PER_SAMPLE * BATCH_PER_GPU / ms_step / peaks["tflops"], 4),
                    "by_kernel": sorted(by_kernel, key=lambda x: -x["ms_per_step"])}

    # ---- generator inference (configs[3]): 30 s utterances, batch 1 per call, round-robin over ranks
    ug = UtteranceGenerator(g, "bf16")
    su_i, sess_i, _ = O.synthetic_batch(1, INFER_FRAMES, seed=7 + rank)
    su_d, sess_d = su_i.to(dev), sess_i.to(dev)
    su_p, sess_p = su_i.pin_memory(), sess_i.pin_memory()
    for _ in range(3):
        ug.generate_graph(su_d, sess_d)
    n_utt = max(10, args.steps)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_utt):
        ug.generate_graph(su_d, sess_d)
    e1.record()
    barrier()
    ms_inf = max_over_ranks(e0.elapsed_time(e1)) / n_utt
    host_out = torch.empty(1, INFER_FRAMES * HOP, CHANNELS).pin_memory()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_utt):
        host_out.copy_(ug.generate_graph(su_p, sess_p), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    e1.record()
    barrier()
    ms_inf_e2e = max_over_ranks(e0.elapsed_time(e1)) / n_utt
    # throughput mode: 4 utterances of equal length per call (same engine, one captured graph per (batch, frames))
    IB = 4
    su_b, sess_b, _ = O.synthetic_batch(IB, INFER_FRAMES, seed=70 + rank)
    su_b, sess_b = su_b.to(dev), sess_b.to(dev)
    for _ in range(3):
        ug.generate_graph(su_b, sess_b)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_utt):
        ug.generate_graph(su_b, sess_b)
    e1.record()
    barrier()
    ms_inf_b = max_over_ranks(e0.elapsed_time(e1)) / n_utt
    inference = {"workload": "generator-only, 1500 unit frames -> 24000x8 EMG samples per utterance, batch 1, bf16 (configs[3])",
                 "batch4": {"utterances_per_s": world * IB / (ms_inf_b / 1e3), "ms_per_call": ms_inf_b,
                            "tensor_frac": round(G_FWD_GFLOP_PER_FRAME * INFER_FRAMES * IB / ms_inf_b / peaks["tflops"], 4)},
                 "utterances_per_s": world / (ms_inf / 1e3), "emg_samples_per_s": world * INFER_FRAMES * HOP / (ms_inf / 1e3),
                 "ms_per_utterance": ms_inf, "e2e_utterances_per_s": world / (ms_inf_e2e / 1e3),
                 "tensor_frac": round(G_FWD_GFLOP_PER_FRAME * INFER_FRAMES / ms_inf / peaks["tflops"], 4)}

    # ---- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(2, 1, budget_s=30.0)
        cpu = {"value": r["value"], "unit": "samples/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]}

    if rank == 0:
        line = {"metric": "GAN train samples/s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config_dict(world), "clocks": clocks,
                "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps), "launches_per_step": int(launches_per_step),
                "roofline": roofline, "cpu_baseline": cpu, "inference": inference,
                "losses_last_step": {k: round(v, 5) for k, v in final_losses.items()}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
