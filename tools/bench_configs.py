"""Side measurements for the BASELINE.json configs that are not the bench.py headline
(config[0] 60-s clip, config[2] MIC log-mel + GCC-PHAT, config[3] sharded scaler action).
Run directly (1 GPU) or under torchrun (N GPUs).  Prints one JSON line per config (rank 0)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import adyolo_b200 as A
from adyolo_b200.features import features_mic_batched

rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def timed(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    return ms


g = torch.Generator(device=dev).manual_seed(rank)
def audio(B, N):
    return (torch.randn((B, N, 4), device=dev, generator=g) * 3000).clamp_(-32768, 32767).to(torch.int16)

out = []
# config[0]: one 60-s clip (latency of the preprocess.py path)
a60 = audio(1, 1_440_000)
ms = timed(lambda: A.features_batched(a60, None), 20, 5)
out.append({"config": "config[0] one 60-s FOA clip, raw log-mel+IV (preprocess.py path)", "ms": ms, "audio_hours_per_s": 60 / 3600 / (ms / 1e3), "n_gpus": 1})
# config[2]: MIC log-mel + GCC-PHAT, 512 x 5-s chunks sharded over the ranks
Bm = 512 // world
am = audio(Bm, 120_000)
ms = timed(lambda: features_mic_batched(am), 3, 1)
out.append({"config": "config[2] MIC log-mel + GCC-PHAT (10ch), 512 x 5-s chunks (fused log-mel+spectra kernel -> GCC-PHAT kernel)", "ms": ms,
            "audio_hours_per_s": world * Bm * 5 / 3600 / (ms / 1e3), "n_gpus": world})
# config[3]: scaler action, resident pool of 60-s clips streamed in batches + all-reduce at the end
pool = audio(8, 1_440_000)
def scaler_pass():
    acc = A.ScalerAccumulator(7, dev)
    for _ in range(4):                                   # 4 batches x 8 clips x 60 s = 32 audio-minutes per rank
        acc.update(A.features_batched(pool, None))
    return acc.result()
ms = timed(scaler_pass, 5, 2)
out.append({"config": "config[3] scaler action: per-(mel,ch) mean/std/max/min, 32 audio-min per rank per pass + all-reduce", "ms": ms,
            "audio_hours_per_s": world * 32 / 60 / (ms / 1e3), "n_gpus": world})
# N1: decode + conn-merge NMS of one 60-s validation clip (600 label frames), sparse detections
if rank == 0:
    import time as _t
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle.nms_torch import YoloPostOracle
    grid = A.labels.GridSpec(12, 5, [45, 45], 0.5)
    lg = torch.randn((1, 600, 2400), device=dev, generator=g) - 3.0
    yv = lg.view(1, 600, 8, 4, 5, 15)
    yv[0, :, 2, 1, :3, 0] += 7; yv[0, :, 2, 1, :3, 3] += 7; yv[0, ::2, 5, 2, :2, 0] += 7; yv[0, ::2, 5, 2, :2, 9] += 7
    ms = timed(lambda: A.yolo_post_batched(lg, grid, 0.5, 0.5, 15.0), 10, 3) if world == 1 else None
    t0 = _t.perf_counter(); ref = YoloPostOracle(device="cpu").clip_output(lg[0].cpu()); cpu_ms = (_t.perf_counter() - t0) * 1e3
    out.append({"config": "N1 decode + conn-merge NMS, one 60-s clip (600 frames)", "ms": ms, "reference_cpu_ms": cpu_ms,
                "detections": sum(len(v) for v in ref.values()), "n_gpus": 1})
if rank == 0:
    for o in out:
        print(json.dumps(o))
if world > 1:
    dist.destroy_process_group()
