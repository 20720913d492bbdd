"""Executed drop-in check, host half (CPU, needs /root/reference): the reference's OWN, unmodified
``train.train_one_epoch`` and ``test.test_epoch`` (src/train.py:44-60, src/test.py:33-60) are run on a tiny
synthetic DCASE tree twice --

  (1) as the reference wires them: ``datasets.Dataset`` + ``collate_fn`` (features and label rows made in
      ``__getitem__``), reference ``ADYOLOloss`` / ``LabelPostProcessor``;
  (2) as INTEGRATION.md wires them: ``adyolo_b200.RawAudioDataset`` + ``collate_raw`` (product code: int16 audio
      and the event table cross the DataLoader boundary), a front-end module ahead of the same encoder, and a
      criterion / post-processor taking the same arguments as the product classes.

There is no GPU in this container and no reference on the GPU box, so in (2) the three device-side classes
are stood in for by oracle-backed subclasses with the product signatures (what they compute on the device is
pinned separately by the -m gpu parity tests; tests/test_gpu_dropin.py runs the product classes through the
same steps on the B200).  Gate: both wirings give the same loss sequence and the same written SELD csv files.
"""
import os
import random

import numpy as np
import pytest
import torch

from conftest import ROOT  # noqa: F401
from oracle import ref_shims

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="/root/reference not present")


class TinyEncoder(torch.nn.Module):
    """(B, 7, T, 64) -> (B, T/4, 160 * (C + 3)): stands for WrapperModel (stock PyTorch, out of scope)."""

    def __init__(self, nb_classes=12):
        super().__init__()
        self.pool = torch.nn.AvgPool2d((4, 8))
        self.lin = torch.nn.Linear(7 * 8, 160 * (nb_classes + 3))

    def forward(self, x):
        y = self.pool(x)                                   # (B, 7, T/4, 8)
        return self.lin(y.permute(0, 2, 1, 3).flatten(2))


def _stand_ins(params, scaler):
    """CPU stand-ins with the signatures of FrontEndModule / WrapperCriterion / LabelPostProcessor."""
    import adyolo_b200 as A
    from oracle import assign_np, features_np as F
    from oracle.loss_torch import ADYOLOlossOracle
    from oracle.nms_torch import YoloPostOracle

    class Front(A.FrontEndModule):
        def __init__(self):
            torch.nn.Module.__init__(self)

        def extract(self, audio_i16):
            f = [F.features_foa_stack(a.numpy(), scaler=scaler) for a in audio_i16]
            return torch.from_numpy(np.stack(f)).float()

        def forward(self, audio):
            return self.extract(audio.to(torch.int16))

    class Crit:
        def __init__(self):
            self.loss_nm, self.oracle = "adyolo", ADYOLOlossOracle(params)

        def __call__(self, output, target):                 # ADYOLOloss.__call__ with the (E, 5) event table
            assert target.dim() == 2 and target.shape[1] == 5 and target.dtype == torch.float32
            rows = assign_np.events_to_rows(target.numpy().astype(np.float64), output.shape[1])
            return self.oracle(output, torch.from_numpy(rows.astype(np.float32)))

    class Post:
        def __init__(self):
            tc = params["train_config"]
            self.o = YoloPostOracle(params["data_config"]["nb_classes"], conf_thresh=tc["conf_thresh"],
                                    clss_thresh=tc["clss_thresh"], unify_thresh=tc["unify_thresh"])

        def postprocess(self, output):                      # LabelPostProcessor.postprocess((1, T, n) logits)
            return self.o.clip_output(output[0])

    return Front(), Crit(), Post()


def _seed():
    random.seed(7); np.random.seed(7); torch.manual_seed(7)


def test_reference_loops_run_unchanged_on_the_raw_audio_wiring(tmp_path, scaler2021):
    from _tiny_tree import build, params_for
    build(str(tmp_path), scaler2021)
    ref_shims.install()
    import datasets as ref_datasets
    import models.loss as ref_loss
    import train as ref_train
    import test as ref_test
    import adyolo_b200 as A
    from torch.utils.data import DataLoader
    params = params_for(tmp_path)
    dev = torch.device("cpu")

    def run(wiring):
        _seed()
        enc = TinyEncoder()
        if wiring == "reference":
            ds = ref_datasets.Dataset(params, "train")
            dl = DataLoader(ds, batch_size=2, shuffle=False, collate_fn=ref_datasets.collate_fn)
            model, crit = enc, ref_loss.ADYOLOloss(params)
            post = ref_datasets.LabelPostProcessor(params)
            ds_t = ref_datasets.Dataset(params, "test", is_valid=True)
            dl_t = DataLoader(ds_t, batch_size=1, shuffle=False, collate_fn=ref_datasets.collate_fn)
        else:
            front, crit, post = _stand_ins(params, scaler2021)
            ds = A.RawAudioDataset(params, "train")
            dl = DataLoader(ds, batch_size=2, shuffle=False, collate_fn=A.collate_raw, num_workers=2)
            model = torch.nn.Sequential(front, enc)
            ds_t = A.RawAudioDataset(params, "test", is_valid=True)
            dl_t = DataLoader(ds_t, batch_size=1, shuffle=False, collate_fn=A.collate_raw)
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        files = list(ds.get_filelist())
        tl = ref_train.train_one_epoch(params, dl, model, opt, crit, dev)          # unmodified reference loop
        out_dir = os.path.join(str(tmp_path), "out_" + wiring)
        vl = ref_test.test_epoch(dl_t, model, crit, post, dev, out_dir)            # unmodified reference loop
        csvs = {f: open(os.path.join(out_dir, f)).read() for f in sorted(os.listdir(out_dir))}
        return files, tl, vl, csvs

    f_ref, tl_ref, vl_ref, csv_ref = run("reference")
    f_our, tl_our, vl_our, csv_our = run("raw-audio")
    assert f_ref == f_our                                   # same epoch sample under the same seed
    assert np.isfinite(tl_ref) and abs(tl_ref - tl_our) <= 1e-5 * abs(tl_ref), (tl_ref, tl_our)
    assert abs(vl_ref - vl_our) <= 1e-5 * abs(vl_ref), (vl_ref, vl_our)
    assert sorted(csv_ref) == sorted(csv_our) and len(csv_ref) == 2
    for k in csv_ref:                                       # same detections (coordinates to 1e-4)
        a = np.array([[float(v) for v in l.split(",")] for l in csv_ref[k].splitlines()]).reshape(-1, 6)
        b = np.array([[float(v) for v in l.split(",")] for l in csv_our[k].splitlines()]).reshape(-1, 6)
        assert a.shape == b.shape and np.array_equal(a[:, :3], b[:, :3]) and np.abs(a - b).max() < 1e-4


def test_raw_dataset_hands_over_what_the_reference_dataset_consumes(tmp_path, scaler2021):
    """Item by item: the int16 clip and the event table of RawAudioDataset are exactly the inputs of the reference's
    feature / label code, and collate_raw + oracle rows == the reference's collate_fn target."""
    from _tiny_tree import build, params_for
    build(str(tmp_path), scaler2021)
    ref_shims.install()
    import datasets as ref_datasets
    import adyolo_b200 as A
    from oracle import assign_np
    params = params_for(tmp_path)
    random.seed(3)
    ref = ref_datasets.Dataset(params, "train")
    random.seed(3)
    ours = A.RawAudioDataset(params, "train")
    assert ref.get_filelist() == ours.get_filelist() and len(ours) == len(ref) == 10
    assert sorted(ref.get_remaining_file()) == sorted(ours.get_remaining_file())
    batch_ref = [ref[i] for i in range(3)]
    batch_our = [ours[i] for i in range(3)]
    _, tgt = ref_datasets.collate_fn(batch_ref)
    audio, events = A.collate_raw(batch_our)
    assert audio.dtype == torch.int16 and audio.shape == (3, 48000, 4) and events.dtype == torch.float32
    rows = assign_np.events_to_rows(events.numpy().astype(np.float64), 20).astype(np.float32)
    assert np.array_equal(rows, tgt.numpy())
    for i in range(3):
        name = ours.get_filelist()[i]
        assert np.array_equal(batch_our[i][0].numpy(), ref.load_wav2npy(os.path.join(ref.wav_pth, name + ".wav")))
