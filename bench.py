#!/usr/bin/env python
"""bench.py — headline benchmark of the AD-YOLO hot path (BASELINE.json metric / config).

metric : audio-hours of 4-ch features per second (whole job, all N GPUs)
workload (N=1, config[1] of BASELINE.json): batch of 256 synthetic 5-s FOA chunks per GPU:
  fused feature front end (log-mel + intensity vectors, standardised) + AD-YOLO label rows +
  `--loss adyolo` forward/backward on synthetic logits of the se-resnet34 output shape
  (256, 50, 2400).  The encoder itself is stock PyTorch and out of scope (not timed).
A "step" = one pass of that hot path over one batch.  Weak scaling: every rank gets its own
256-clip batch, no data-path collective (clips are independent).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`--impl reference` times the reference's CPU implementation of the same path (the numpy/torch
oracle port of librosa+reference code; librosa itself is not installable here) on the box's host
cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CLIP_S = 5.0
SR = 24000
N_SAMPLES = int(CLIP_S * SR)
T_LABEL = 50
NB_CLASSES = 12
BATCH = 256
BYTES_PER_AUDIO_S = 24000 * 4 * 2 + 40 * 64 * 7 * 4          # 263 680 B
WORKLOAD = "config[1]: 256 x 5-s FOA chunks/GPU: fused log-mel+IV front end + adyolo label rows + adyolo loss fwd/bwd"


# ------------------------------------------------------------------------------------------------ synthetic data
def synth_audio(rng, n_clips):
    return np.clip(rng.standard_normal((n_clips, N_SAMPLES, 4), dtype=np.float32) * 3000.0, -32768, 32767).astype(np.int16)


def synth_events(rng, n_clips):
    """0-3 events per label frame (incl. same-class overlaps), integer degrees incl. poles / +-180."""
    n = rng.integers(0, 4, size=(n_clips, T_LABEL))
    b, t = np.nonzero(n)
    reps = n[b, t]
    b, t = np.repeat(b, reps), np.repeat(t, reps)
    E = len(b)
    same = rng.integers(0, NB_CLASSES, size=(n_clips, T_LABEL))[b, t]
    cls = np.where(rng.random(E) < 0.5, same, rng.integers(0, NB_CLASSES, size=E))
    az = rng.integers(-180, 181, E).astype(np.float64)
    el = rng.integers(-90, 91, E).astype(np.float64)
    return np.stack([b, t, cls, az, el], 1).astype(np.float64)


def load_scaler():
    z = np.load(os.path.join(ROOT, "tests", "golden", "scaler_DCASE2021.npz"))
    return {"MEL": {k: z[f"MEL_{k}"] for k in ("mean", "std")}, "IV": {k: z[f"IV_{k}"] for k in ("mean", "std")}}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def _cpu_features_one(args):
    clip, scaler = args
    from oracle import features_np as F
    return F.features_foa_stack(clip, scaler=scaler).astype(np.float32)


def cpu_reference_step(pool, clips, events, scaler, logit):
    """One step of the same hot path on the host: oracle features (one clip per worker process, like
    the reference's DataLoader workers), numpy label rows, torch-CPU ADYOLOloss fwd+bwd."""
    import torch
    from oracle import assign_np
    from oracle.loss_torch import ADYOLOlossOracle, default_params
    feats = pool.map(_cpu_features_one, [(c, scaler) for c in clips])
    rows = torch.from_numpy(assign_np.events_to_rows(events, T_LABEL).astype(np.float32))
    crit = ADYOLOlossOracle(default_params(NB_CLASSES, "cpu"))
    logit.grad = None
    loss = crit(logit, rows)
    loss.backward()
    return feats, float(loss)


def run_cpu_baseline(steps, warmup, sample_clips=None):
    """Returns (audio_hours_per_s, cores, sample description)."""
    import multiprocessing as mp
    import torch
    cores = os.cpu_count() or 1
    n = sample_clips or min(max(2 * cores, 16), 128)
    rng = np.random.default_rng(123)
    clips = synth_audio(rng, n)
    events = synth_events(rng, n)
    scaler = load_scaler()
    logit = torch.randn(n, T_LABEL, 160 * (NB_CLASSES + 3), generator=torch.Generator().manual_seed(0)).requires_grad_(True)
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(warmup):
            cpu_reference_step(pool, clips[:cores], events[events[:, 0] < cores], scaler, logit[:cores].detach().requires_grad_(True))
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_reference_step(pool, clips, events, scaler, logit)
        dt = time.perf_counter() - t0
    hours = steps * n * CLIP_S / 3600.0
    return hours / dt, cores, f"{n} x 5-s clips/step x {steps} steps, {cores} worker processes (numpy f64 oracle port) + torch-CPU loss", dt / steps


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region.  In-process NVML (pynvml, the data
    source nvidia-smi prints) every 10 ms, so that even a ~100 ms timed region yields a real median;
    falls back to spawning nvidia-smi (one sample per ~0.3 s) when pynvml is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, pci_bus_id=None):
        super().__init__(daemon=True)
        self.gpu, self.samples, self._stop_evt = gpu_index, [], threading.Event()
        self.nv = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = (pynvml.nvmlDeviceGetHandleByPciBusId(pci_bus_id.encode()) if pci_bus_id
                           else pynvml.nvmlDeviceGetHandleByIndex(gpu_index))
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nv = pynvml
        except Exception:
            self.nv = self.handle = None

    def _sample_nvml(self):
        nv = self.nv
        sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        r = int(get(self.handle))
        flag = lambda bit: "Active" if r & bit else "Not Active"   # noqa: E731  (NVML reason bits)
        self.samples.append([str(self.gpu), str(sm), str(self.max_sm), "", "", flag(0x8), flag(0x40), flag(0x20), flag(0x4)])

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self.nv is not None:
                    self._sample_nvml()
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.01 if self.nv is not None else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(s[1]) for s in self.samples if len(s) > 2 and s[1].replace(".", "").isdigit()]
        mx = [float(s[2]) for s in self.samples if len(s) > 2 and s[2].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            if len(s) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples), "source": "nvml" if self.nv is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------ ours
def main_ours(args):
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))

    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        # before CUDA is initialised in this process (the pool forks)
        v, cores, sample, _ = run_cpu_baseline(steps=1, warmup=1)
        cpu_base = {"value": v, "unit": "audio-hours/s", "cores": cores, "kind": "port", "sample": sample}

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    if rank == 0 or world == 1:
        pass
    import adyolo_b200 as A
    from adyolo_b200.features import _scaler_to_device
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # node-local pinned staging buffers: without this the e2e number stops scaling at 4 ranks
    numa = A.bind_host_to_device(dev)

    rng = np.random.default_rng(1234 + rank)
    audio_h = torch.from_numpy(synth_audio(rng, BATCH)).pin_memory()
    events_h = torch.from_numpy(synth_events(rng, BATCH)).pin_memory()
    audio_d = audio_h.to(dev)
    events_d = events_h.to(dev)
    scaler_dev = _scaler_to_device(load_scaler(), ("MEL", "IV"), dev)
    params = {"args": {"device": str(dev), "loss": "adyolo"}, "data_config": {"nb_classes": NB_CLASSES},
              "train_config": {"grid_size": [45, 45], "nb_anchors": 5, "g_overlap": 0.5, "train_unify": [45., 25., 10.],
                               "loss_gains": {"angular_gain": 5., "object_gain": 1., "nonobj_gain": 5., "class_gain": 3.}}}
    crit = A.ADYOLOloss(params)
    grid = crit.grid
    logit = torch.randn((BATCH, T_LABEL, 160 * (NB_CLASSES + 3)), device=dev,
                        generator=torch.Generator(device=dev).manual_seed(rank)).requires_grad_(True)
    feat_buf = torch.empty((BATCH, 7, N_SAMPLES // 600, 64), dtype=torch.float32, device=dev)
    fe_events = []

    def step(audio, events, record=False):
        A.features_batched(audio, scaler_dev, out=feat_buf, timing_events=fe_events if record else None)
        rows = A.label_rows_batched(events, T_LABEL, grid, max_rows=4 * events.shape[0])   # no host sync
        logit.grad = None
        loss = crit(logit, rows)
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value)
    for _ in range(args.warmup):
        step(audio_d, events_d)
    barrier()
    props = torch.cuda.get_device_properties(dev)
    try:
        pci = "%08x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
    except AttributeError:
        pci = None
    sampler = ClockSampler(local_rank, pci)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(audio_d, events_d, record=True)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    fe_ms = float(np.mean([a.elapsed_time(b) for a, b in fe_events])) if fe_events else None

    # ---- end-to-end through the public API with host buffers: every step's int16 batch and event
    # table are copied from pinned host memory (double-buffered on a side stream, so the copy of
    # step i+1 overlaps the kernels of step i) and the step's loss is read back to the host
    pipe = A.HostBatchPipeline(dev)

    def e2e_loop(n):
        pipe.submit(audio_h, events_h)
        last = None
        for i in range(n):
            a_d, e_d = pipe.get()
            if i + 1 < n:
                pipe.submit(audio_h, events_h)
            loss = step(a_d, e_d)
            pipe.release()
            last = loss.item()                          # device -> host read of the step's result
        return last

    e2e_loop(max(2, args.warmup // 2))
    barrier()
    x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x0.record()
    lv = e2e_loop(args.steps)
    x1.record()
    barrier()
    e2e_ms = x0.elapsed_time(x1)
    clocks = sampler.stop()

    if world > 1:
        t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = t.tolist()
    hours_per_step = world * BATCH * CLIP_S / 3600.0
    value = hours_per_step * args.steps / (ms / 1e3)
    e2e_value = hours_per_step * args.steps / (e2e_ms / 1e3)

    peaks, peak_src = None, "fallback"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        peak, peak_src = float(peaks["hbm_gbs"]), "measured"
    except Exception:
        peak = 6650.0
    roof = None
    if fe_ms:
        alg_bytes = BATCH * CLIP_S * BYTES_PER_AUDIO_S
        ach = alg_bytes / (fe_ms / 1e3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "frontend_traffic.json")) as f:
                traffic = json.load(f)["dram_bytes_total"]      # ncu --set full capture of the same launch shape
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": "frontend_foa_kernel", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": traffic, "peak_source": peak_src, "kernel_ms": fe_ms,
                "algorithmic_bytes_per_launch": alg_bytes,
                # the north star also asks for the FP32 (CUDA-core) roofline: algorithmic 6.5 MFLOP per audio-second
                #  against 148 SM x 128 lanes x 2 FLOP x the sampled SM clock
                "fp32": {"achieved": BATCH * CLIP_S * 6.5e6 / (fe_ms / 1e3) / 1e12,
                         "peak": 148 * 128 * 2 * (clocks.get("sm_mhz") or 1965.0) * 1e6 / 1e12, "unit": "TFLOP/s"},
                "note": "FP32-issue-bound kernel (25 FLOP/B vs ~11 FLOP/B ridge); see DESIGN.md / profiles/"}
        roof["fp32"]["frac"] = roof["fp32"]["achieved"] / roof["fp32"]["peak"]
    if rank == 0:
        line = {"metric": "audio-hours of 4-ch features/sec", "value": value, "unit": "audio-hours/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "clips_per_gpu": BATCH, "clip_s": CLIP_S, "nb_classes": NB_CLASSES,
                           "l2": "inputs larger than L2 (246 MB int16 audio + 123 MB logits per step)",
                           "parallelism": f"clip-sharded x{world}, no collective"},
                "e2e": {"value": e2e_value, "unit": "audio-hours/s", "ms_per_step": e2e_ms / args.steps,
                        "h2d_bytes_per_step": int(audio_h.numel() * 2 + events_h.numel() * 8), "d2h_bytes_per_step": 4 + 8},
                "gpu_launches": args.steps * 10, "roofline": roof, "cpu_baseline": cpu_base, "clocks": clocks, "host_numa": numa,
                "loss": lv}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main_reference(args):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    v, cores, sample, s_per_step = run_cpu_baseline(steps=args.steps, warmup=min(args.warmup, 1))
    line = {"impl": "reference", "metric": "audio-hours of 4-ch features/sec", "value": v, "unit": "audio-hours/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD + " (bounded sample on host cores)"},
            "cpu_baseline": {"value": v, "unit": "audio-hours/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "audio-hours/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        main_reference(a)
    else:
        main_ours(a)
