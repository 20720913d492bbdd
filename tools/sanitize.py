"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import adyolo_b200 as A
from adyolo_b200.features import features_mic_batched
rng = np.random.default_rng(0)
clips = torch.from_numpy(np.clip(rng.standard_normal((5, 600 * 10 + 77, 4)) * 3000, -32768, 32767).astype(np.int16)).cuda()
f = A.features_batched(clips)
m = features_mic_batched(clips[:2])
grid = A.labels.GridSpec(12, 5, [45, 45], 0.5)
ev = np.array([[b, t, rng.integers(0, 12), rng.integers(-180, 181), rng.integers(-90, 90)] for b in range(5) for t in range(10) for _ in range(2)], dtype=np.float64)
rows = A.label_rows_batched(torch.from_numpy(ev).cuda(), 10, grid)
drows = A.label_rows_batched(torch.from_numpy(ev).cuda(), 10, grid, max_rows=4 * len(ev))
p = {"args": {"device": "cuda:0", "loss": "adyolo"}, "data_config": {"nb_classes": 12},
     "train_config": {"grid_size": [45, 45], "nb_anchors": 5, "g_overlap": 0.5, "train_unify": [45., 25., 10.],
                      "loss_gains": {"angular_gain": 5., "object_gain": 1., "nonobj_gain": 5., "class_gain": 3.}}}
crit = A.ADYOLOloss(p)
for tgt in (rows, drows):
    logit = torch.randn(5, 10, 2400, device="cuda", requires_grad=True)
    loss = crit(logit, tgt); loss.backward()
D, mk, am = A.adyolo_assign(logit, rows, grid)
acc = A.ScalerAccumulator(7); acc.update(A.features_batched(clips, None)); r = acc.result()
torch.cuda.synchronize()
print("sanitize pass ok", float(loss), f.shape, m.shape)
