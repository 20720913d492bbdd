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
(crit(logit, rows) * 0.5).backward()                       # grad_scale path with a non-unit upstream gradient
sa = A.SpecAug({"aug_config": {"spec_augment": True, "spec_augment_thresh": 1.0, "spec_augment_time_mask_param": 40,
                               "spec_augment_freq_mask_param": 40}}, is_valid=False)
f = sa.augment_batched(f)
dets = A.yolo_post_batched(logit.detach()[:1], grid, 0.5, 0.5, 15.0)
acc = A.ScalerAccumulator(7); acc.update(A.features_batched(clips, None)); r = acc.result()
torch.cuda.synchronize()
print("sanitize pass ok", float(loss), f.shape, m.shape)
