"""config[1] step with the front end and the label / loss chain on two CUDA streams (they are independent: features come from
the audio, the loss from the encoder's logits), against the same step on one stream.  ADYOLO_LIB selects a variant build."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import adyolo_b200 as A
from oracle.loss_torch import default_params
import bench

B, T, C = 256, 50, 12
rng = np.random.default_rng(0)
ev = torch.from_numpy(bench.synth_events(rng, B)).cuda()
grid = A.labels.GridSpec(C, 5, [45, 45], 0.5)
crit = A.ADYOLOloss(default_params(C, "cuda:0"))
logit = torch.randn((B, T, 2400), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)).requires_grad_(True)
g = torch.Generator(device="cuda").manual_seed(0)
audio = (torch.randn((B, 120000, 4), device="cuda", generator=g) * 3000).clamp_(-32768, 32767).to(torch.int16)
feat = torch.empty((B, 7, 200, 64), device="cuda")
sF, sL = torch.cuda.Stream(), torch.cuda.Stream()


def loss_side():
    rows = A.label_rows_batched(ev, T, grid, max_rows=4 * ev.shape[0])
    logit.grad = None
    loss = crit(logit, rows)
    loss.backward()
    return loss


def step_serial():
    A.features_batched(audio, None, out=feat)
    return loss_side()


def step_two_streams():
    with torch.cuda.stream(sF):
        A.features_batched(audio, None, out=feat)
    with torch.cuda.stream(sL):
        return loss_side()


def timed(fn, n=40):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    cur = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sF.wait_stream(cur); sL.wait_stream(cur)
    e0.record()
    sF.wait_event(e0); sL.wait_event(e0)
    for _ in range(n):
        loss = fn()
    cur.wait_stream(sF); cur.wait_stream(sL)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, float(loss)


for name, fn in (("one stream", step_serial), ("two streams", step_two_streams)):
    ms, loss = timed(fn)
    print(f"{name:12s} {ms:.4f} ms/step  ({B * 5 / 3600 / (ms / 1e3):.0f} audio-h/s)  loss {loss:.6f}  lib {os.environ.get('ADYOLO_LIB', 'default')}")
