"""Timing driver for the MIC path (config[2]): 512 x 5-s clips -> (512, 10, 200, 64)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from adyolo_b200.features import features_mic_batched
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
g = torch.Generator(device="cuda").manual_seed(0)
audio = (torch.randn((B, 120000, 4), device="cuda", generator=g) * 3000).clamp_(-32768, 32767).to(torch.int16)
for _ in range(3):
    features_mic_batched(audio)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    features_mic_batched(audio)
e1.record(); torch.cuda.synchronize()
print("features_mic_batched ms/call:", e0.elapsed_time(e1) / 10)
