"""features_batched timings over the clip shapes of BASELINE.json's side configs (ms per call)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adyolo_b200 as A
g = torch.Generator(device="cuda").manual_seed(0)
for B, secs in ((1, 60), (8, 60), (16, 20), (256, 5)):
    audio = (torch.randn((B, secs * 24000, 4), device="cuda", generator=g) * 3000).clamp_(-32768, 32767).to(torch.int16)
    out = torch.empty((B, 7, secs * 40, 64), device="cuda")
    for _ in range(3):
        A.features_batched(audio, None, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        A.features_batched(audio, None, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{B:4d} x {secs:2d} s: {ms:.4f} ms  ({B * secs / 3600 / (ms / 1e3):.0f} audio-h/s)")
