"""Timing driver for the label -> assignment -> loss (forward + backward) half of the step at config[1]'s shape
(B=256 x 50 label frames, C=12: 2.05 M anchors, 123 MB of logits).  ADYOLO_LIB selects a variant build."""
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


def step():
    rows = A.label_rows_batched(ev, T, grid, max_rows=4 * ev.shape[0])
    loss = crit(logit, rows)
    logit.grad = None
    loss.backward()
    return loss


for _ in range(5):
    step()
torch.cuda.synchronize()
n = 30
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    loss = step()
e1.record(); torch.cuda.synchronize()
print("labels + loss fwd/bwd ms/step: %.4f   loss %.6f   lib %s" % (e0.elapsed_time(e1) / n, loss.item(), os.environ.get("ADYOLO_LIB", "default")))
