"""Timing driver for the scaler action (config[3] shape: 8 x 60-s clips per update, 8 updates per pass)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adyolo_b200 as A
g = torch.Generator(device="cuda").manual_seed(0)
pool = (torch.randn((8, 1_440_000, 4), device="cuda", generator=g) * 3000).clamp_(-32768, 32767).to(torch.int16)
def one_pass():
    acc = A.ScalerAccumulator(7, "cuda")
    for _ in range(8):
        acc.update_from_audio(pool)
    return acc
for _ in range(2):
    one_pass()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    one_pass()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("scaler pass ms:", ms, "audio-h/s:", (64 / 60) / (ms / 1e3))
