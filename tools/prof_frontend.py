"""Tiny driver for ncu captures of the front-end kernel (run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import adyolo_b200 as A
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator(device="cuda").manual_seed(0)
audio = (torch.randn((B, 120000, 4), device="cuda", generator=g) * 3000).clamp_(-32768, 32767).to(torch.int16)
out = torch.empty((B, 7, 200, 64), device="cuda")
scaler = None
if len(sys.argv) > 2 and sys.argv[2] == "scaler":      # standardisation constants as in bench.py (tests/golden)
    import numpy as np
    from adyolo_b200.features import _scaler_to_device
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scaler_DCASE2021.npz"))
    scaler = _scaler_to_device({g: {k: z[f"{g}_{k}"] for k in ("mean", "std")} for g in ("MEL", "IV")}, ("MEL", "IV"), torch.device("cuda"))
for _ in range(3):
    A.features_batched(audio, scaler, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    A.features_batched(audio, scaler, out=out)
e1.record(); torch.cuda.synchronize()
print("features_batched ms/call:", e0.elapsed_time(e1) / 10)
