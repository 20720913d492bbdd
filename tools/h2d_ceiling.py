"""Raw host->device copy ceiling of the box: every rank copies the 246.5 MB int16 batch of
config[1] from pinned host memory at the same time, as (a) one copy, (b) 4 chunks on one stream, (c) 4 chunks over two
streams.  Launch: python tools/h2d_ceiling.py (1 GPU) or torchrun --nproc-per-node N tools/h2d_ceiling.py.
Prints one line per variant: GB/s per GPU (slowest rank) and aggregate."""
import os, sys
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
nbytes = 256 * 120000 * 4 * 2 + 765728
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
host.random_(0, 255)
dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def one_copy():
    dst.copy_(host, non_blocking=True)


def chunks(streams, n=4):
    step = (nbytes + n - 1) // n
    cur = torch.cuda.current_stream()
    for s in streams:
        s.wait_stream(cur)
    for i in range(n):
        with torch.cuda.stream(streams[i % len(streams)]):
            dst[i * step:(i + 1) * step].copy_(host[i * step:(i + 1) * step], non_blocking=True)
    for s in streams:
        cur.wait_stream(s)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


for name, fn in (("one copy", one_copy), ("4 chunks, 1 stream", lambda: chunks([s1])), ("4 chunks, 2 streams", lambda: chunks([s1, s2]))):
    ms = timed(fn)
    if rank == 0:
        gbs = nbytes / (ms / 1e3) / 1e9
        print(f"ranks {world}  {name:20s} {ms:7.3f} ms  {gbs:6.1f} GB/s per GPU  {gbs * world:7.1f} GB/s aggregate", flush=True)
if world > 1:
    dist.destroy_process_group()
