"""Generate tests/golden/*.npz by executing the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the authoring container only (needs /root/reference):  ``python oracle/make_golden.py``

Fixtures written (all small, committed):

* ``scaler_DCASE20xx.npz``  — the reference's shipped ``data/DCASE20xx_SELD/scaler_wts.pkl``
  statistics (real fixtures, inputs to standardisation), converted to npz.
* ``features_foa.npz``      — int16 clips + the outputs of the reference's own
  ``FeatureLabelProcessor.get_feature`` / ``utility.audio2stft|stft2melscale|stft2iv`` executed
  over the librosa-0.8.1 restatement (see oracle/ref_shims.py; librosa itself is absent here).
* ``assign_cells.npz``      — reference ``get_yolo_label`` + ``collate_fn`` on a label dict, and
  the responsible-cell bitmask of every integer (azi, ele) on the sphere plus boundary values.
* ``loss_ref.npz``          — reference ``ADYOLOloss`` (device='cpu', the only device it runs on
  under torch>=2) loss value and d loss / d logit on seeded inputs that include the
  pole/threshold cases of SURVEY F9; D from the reference's own
  ``distance_between_polar_coordinates``.
"""
from __future__ import annotations

import copy
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def synth_clip(seed: int, seconds: float, kind: str) -> np.ndarray:
    """Synthetic (N,4) int16 FOA clip (channel order W,Y,Z,X as in the DCASE foa_dev wavs).

    'noise'  : ~-20 dBFS white noise, independent per channel.
    'bursts' : three point sources (tone bursts) encoded to FOA by their direction, a -60 dBFS
               sensor-noise floor and 0.5 s of digital silence (exercises amin, the 1e-8
               offset and the top_db clamp).
    'harsh'  : per-channel independent full-scale tone bursts over a 2-LSB dither (channels up
               to 75 dB apart inside one frame: probes the FP32 dynamic-range floor of a GPU
               front end; not a realistic FOA signal)."""
    rng = np.random.default_rng(seed)
    n = int(24000 * seconds)
    t = np.arange(n) / 24000.0
    if kind == "noise":
        x = rng.standard_normal((n, 4)) * 3000.0
    elif kind == "bursts":
        x = rng.standard_normal((n, 4)) * 30.0
        for _ in range(3):
            f = rng.uniform(150, 9000)
            t0 = rng.uniform(0, seconds * 0.5)
            env = ((t > t0) & (t < t0 + 0.3 * seconds)).astype(np.float64)
            s = env * rng.uniform(2000, 9000) * np.sin(2 * np.pi * f * t + rng.uniform(0, 6.28))
            az, el = np.deg2rad(rng.uniform(-180, 180)), np.deg2rad(rng.uniform(-60, 60))
            gains = np.array([1.0, np.sin(az) * np.cos(el), np.sin(el), np.cos(az) * np.cos(el)])
            x += s[:, None] * gains[None, :]
        s0 = int(n * 0.75)
        x[s0:s0 + 12000] = 0.0  # digital silence
    else:
        x = np.zeros((n, 4))
        for c in range(4):
            for _ in range(3):
                f = rng.uniform(100, 11000)
                t0 = rng.uniform(0, seconds * 0.6)
                env = ((t > t0) & (t < t0 + 0.25 * seconds)).astype(np.float64)
                x[:, c] += env * rng.uniform(500, 12000) * np.sin(2 * np.pi * f * t + rng.uniform(0, 6.28))
        x += rng.standard_normal((n, 4)) * 2.0
    return np.clip(np.round(x), -32768, 32767).astype(np.int16)


def synth_labels(seed: int, nb_label_frames: int, nb_classes: int, max_events=3):
    """label dict {frame: [[cls, src, azi, ele], ...]} with integer degrees incl. F9 cases."""
    rng = np.random.default_rng(seed)
    special_el = [90, -90, 65, -65, 45, -45, 80, -80, 67, -68, 0, 22, -23]
    special_az = [180, -180, 179, -179, 0, 135, -135, 157, -158]
    label = {}
    for f in range(nb_label_frames + 3):  # a few frames past the end (must be dropped)
        k = int(rng.integers(0, max_events + 1))
        if k == 0:
            continue
        evs = []
        same_cls = int(rng.integers(0, nb_classes))
        for e in range(k):
            cls = same_cls if rng.random() < 0.5 else int(rng.integers(0, nb_classes))
            az = float(rng.choice(special_az)) if rng.random() < 0.3 else float(rng.integers(-180, 181))
            el = float(rng.choice(special_el)) if rng.random() < 0.3 else float(rng.integers(-90, 91))
            evs.append([cls, e, az, el])
        label[f] = evs
    return label


def gen_overlap_cells():
    """Cell masks of every integer (azi, ele) for g_overlap values that are NOT exact in float32
    (0.2, 0.4): the unmodified reference get_yolo_label (datasets.py:219-238,457-482) with
    train_config.g_overlap changed.  Pins the float64 grid bounds of the C ABI."""
    ref_shims.install()
    import datasets as ref_datasets
    out = {}
    az = np.arange(-180, 181, dtype=np.float64)
    el = np.arange(-90, 91, dtype=np.float64)
    AZ, EL = [a.ravel() for a in np.meshgrid(az, el, indexing="ij")]
    for ov in (0.2, 0.4):
        params = ref_shims.ref_params(12)
        params["train_config"]["g_overlap"] = ov
        flp = ref_datasets.FeatureLabelProcessor(params)
        masks = np.zeros(len(AZ), dtype=np.uint32)
        for i, (a, e) in enumerate(zip(AZ, EL)):
            m = 0
            for row in flp.get_yolo_label({0: [[0, 0, float(a), float(e)]]}, 1):
                m |= 1 << (int(row[1]) * 4 + int(row[2]))
            masks[i] = m
        out["mask_%02d" % int(round(ov * 10))] = masks
    np.savez_compressed(os.path.join(GOLD, "assign_cells_overlap.npz"), sweep_az=AZ, sweep_el=EL, **out)


def main():
    ref_shims.install()
    import torch
    import datasets as ref_datasets
    import models.loss as ref_loss
    import utils.utility as ref_utility
    from oracle import features_np as F
    from oracle.loss_torch import ADYOLOlossOracle

    os.makedirs(GOLD, exist_ok=True)

    # ---- scaler fixtures
    for year in (2020, 2021, 2022):
        with open(f"/root/reference/data/DCASE{year}_SELD/scaler_wts.pkl", "rb") as f:
            sc = pickle.load(f)
        np.savez(os.path.join(GOLD, f"scaler_DCASE{year}.npz"),
                 **{f"{k}_{kk}": vv for k, v in sc.items() for kk, vv in v.items()})

    # ---- features through the reference's own class / functions
    params = ref_shims.ref_params(12)
    flp = ref_datasets.FeatureLabelProcessor(params)
    out = {}
    for name, kind, secs, seed in (("noise", "noise", 2.0, 1), ("bursts", "bursts", 3.0, 2), ("harsh", "harsh", 1.5, 3)):
        clip = synth_clip(seed, secs, kind)
        audio = clip / 32768.0 + 1e-8                                   # datasets.py:147
        (MEL, IV), nlf = flp.get_feature(audio)                          # datasets.py:281-292
        T = int(len(audio) / 600.0)
        spec = ref_utility.audio2stft(audio, T, 1200, 600, 1200, "han")  # utility.py:142
        mel_raw = ref_utility.stft2melscale(spec, 24000, 1200, 64)       # utility.py:168
        iv_raw = ref_utility.stft2iv(spec, 24000, 1200, 64)              # utility.py:194
        out[f"{name}_audio"] = clip
        out[f"{name}_MEL"] = MEL
        out[f"{name}_IV"] = IV
        out[f"{name}_mel_raw"] = mel_raw
        out[f"{name}_iv_raw"] = iv_raw
        out[f"{name}_nlf"] = np.int64(nlf)
        out[f"{name}_spec_probe"] = spec[::17, ::13, :].copy()          # sparse probe of the STFT
    np.savez_compressed(os.path.join(GOLD, "features_foa.npz"), **out)

    # ---- grid-cell assignment
    nlf = 50
    label = synth_labels(3, nlf, 12)
    label_in = copy.deepcopy(label)
    rows = flp.get_yolo_label(copy.deepcopy(label), nlf)                 # datasets.py:457-482
    feat_dummy = torch.zeros(1)
    batch = [(feat_dummy, flp.get_yolo_label(copy.deepcopy(label), nlf)),
             (feat_dummy, []),
             (feat_dummy, flp.get_yolo_label(synth_labels(4, nlf, 12), nlf))]
    _, tgt = ref_datasets.collate_fn(batch)                              # datasets.py:164-184
    # sphere sweep: every integer (azi, ele) + multiples of 22.5 and their float neighbours
    az = np.concatenate([np.arange(-180, 181, dtype=np.float64), np.arange(-202.5, 203, 22.5),
                         np.nextafter(np.arange(-180, 181, 22.5), 1e9), np.nextafter(np.arange(-180, 181, 22.5), -1e9)])
    el = np.concatenate([np.arange(-90, 91, dtype=np.float64), np.arange(-90, 91, 22.5),
                         np.nextafter(np.arange(-90, 91, 22.5), 1e9), np.nextafter(np.arange(-90, 91, 22.5), -1e9)])
    AZ, EL = np.meshgrid(az, el, indexing="ij")
    AZ, EL = AZ.ravel(), EL.ravel()
    masks = np.zeros(len(AZ), dtype=np.uint32)
    for i, (a, e) in enumerate(zip(AZ, EL)):
        r = flp.get_yolo_label({0: [[0, 0, float(a), float(e)]]}, 1)
        m = 0
        for row in r:
            m |= 1 << (int(row[1]) * 4 + int(row[2]))
        masks[i] = m
    ev_frames, ev_rows = [], []
    for fr, evs in label_in.items():
        for ev in evs:
            ev_frames.append(fr)
            ev_rows.append(ev)
    np.savez_compressed(os.path.join(GOLD, "assign_cells.npz"),
                        label_frames=np.asarray(ev_frames, np.int64), label_events=np.asarray(ev_rows, np.float64),
                        nlf=np.int64(nlf), rows=np.asarray(rows, np.float64), collate_target=tgt.numpy(),
                        label2_frames=np.asarray([fr for fr, evs in synth_labels(4, nlf, 12).items() for _ in evs], np.int64),
                        label2_events=np.asarray([ev for fr, evs in synth_labels(4, nlf, 12).items() for ev in evs], np.float64),
                        sweep_az=AZ, sweep_el=EL, sweep_mask=masks)

    # ---- loss (reference class on CPU)
    lo = {}
    for C, seed in ((12, 10), (13, 11), (14, 12)):
        p = ref_shims.ref_params(C)
        crit = ref_loss.ADYOLOloss(p)
        orc = ADYOLOlossOracle(p)
        B, T = 3, 12
        flpC = ref_datasets.FeatureLabelProcessor(p)
        batch = [(feat_dummy, flpC.get_yolo_label(synth_labels(seed * 7 + b, T, C), T)) for b in range(B)]
        _, target = ref_datasets.collate_fn(batch)
        g = torch.Generator().manual_seed(seed)
        logit = torch.randn(B, T, 160 * (C + 3), generator=g)
        logit.requires_grad_(True)
        loss = crit(logit, target)                                       # loss.py:189-251
        loss.backward()
        with torch.no_grad():
            dec = orc.decode(logit)
            bi, ti, gi, gj = (target[:, k].long() for k in range(4))
            D = crit.distance_between_polar_coordinates(dec[bi, ti, gi, gj][..., -2:],
                                                        target[:, None, -2:].repeat(1, 5, 1))  # loss.py:182-187
        lo[f"C{C}_logit"] = logit.detach().numpy()
        lo[f"C{C}_target"] = target.numpy()
        lo[f"C{C}_loss"] = loss.detach().numpy()
        lo[f"C{C}_grad"] = logit.grad.numpy()
        lo[f"C{C}_D"] = D.numpy()
    np.savez_compressed(os.path.join(GOLD, "loss_ref.npz"), **lo)

    # ---- rotation augmentation (reference class, all 16 combinations)
    import utils.augmentations as ref_aug
    pa = ref_shims.ref_params(12)
    pa["aug_config"]["rotation_augment"] = True
    rot = ref_aug.RotationAug(pa, is_valid=False)
    rng = np.random.default_rng(77)
    snippet = rng.integers(-32768, 32768, size=(96, 4)).astype(np.int16)
    snippet[0] = [-32768, -32768, 32767, -32768]                      # int16 wrap case of -1 * -32768
    lab = synth_labels(5, 20, 12)
    ro = {"snippet": snippet, "label_frames": np.asarray([fr for fr, evs in lab.items() for _ in evs], np.int64),
          "label_events": np.asarray([ev for fr, evs in lab.items() for ev in evs], np.float64)}
    for c in range(16):
        a2, l2 = rot._rotate(snippet.copy(), copy.deepcopy(lab), comb_no=c)   # augmentations.py:84-111
        ro[f"audio_{c}"] = np.asarray(a2)
        ro[f"events_{c}"] = np.asarray([ev for fr, evs in l2.items() for ev in evs], np.float64)
        ro[f"rows_{c}"] = np.asarray(flp.get_yolo_label(copy.deepcopy(l2), 20), np.float64)
    # draw sequence of the reference class (comb_no=None -> int(random.uniform(0, 16)), augmentations.py:90-93),
    # identified by matching the augmented audio against the 16 known outputs
    import random as _random
    _random.seed(21)
    draws = []
    for _ in range(48):
        a2, _l2 = rot.augment(snippet.copy(), copy.deepcopy(lab))
        hit = [c for c in range(16) if np.array_equal(np.asarray(a2), ro[f"audio_{c}"])]
        assert len(hit) == 1
        draws.append(hit[0])
    ro["draw_seed"] = np.int64(21)
    ro["draws"] = np.asarray(draws, np.int64)
    np.savez_compressed(os.path.join(GOLD, "rotation.npz"), **ro)
    # ---- SpecAug (reference class executed as-is on (C,T,F) tensors of ones; seeded python + torch RNG)
    import random
    ps = ref_shims.ref_params(12)
    ps["aug_config"].update({"spec_augment": True, "spec_augment_thresh": 0.5,
                             "spec_augment_time_mask_param": 40, "spec_augment_freq_mask_param": 40})
    sa = ref_aug.SpecAug(ps, is_valid=False)
    random.seed(11)
    torch.manual_seed(11)
    sm = np.zeros((24, 2, 200, 64), np.uint8)
    for clip in range(24):
        for gi, ch in enumerate((4, 3)):                       # datasets.py:158-160: MEL group, then IV group
            y = sa.augment(torch.ones(ch, 200, 64))             # augmentations.py:28-33
            assert bool((y == y[0:1]).all())
            sm[clip, gi] = (y[0] == 0).numpy()
    np.savez_compressed(os.path.join(GOLD, "specaug.npz"), masked=sm, seed=np.int64(11), thresh=np.float64(0.5),
                        time_mask_param=np.int64(40), freq_mask_param=np.int64(40))
    # ---- chunking action + epoch sampler (reference functions executed as-is)
    import preprocess as ref_pre
    cp = {"sr": 24000, "chunk_window_s": 20, "chunk_stride_s": 1, "label_hop_len_s": 0.1}
    rng = np.random.default_rng(31)
    n = int(24000 * 26.3)
    ca = rng.integers(-3000, 3000, size=(n, 4)).astype(np.int16)
    clab = synth_labels(9, 263, 12)
    chunks = ref_pre.chunk_instance(ca.copy(), copy.deepcopy(clab), cp)          # preprocess.py:13-48
    ck = {"audio_seed": np.int64(31), "n_samples": np.int64(n), "n_chunks": np.int64(len(chunks)),
          "label_frames": np.asarray([fr for fr, evs in clab.items() for _ in evs], np.int64),
          "label_events": np.asarray([ev for fr, evs in clab.items() for ev in evs], np.float64),
          "chunk_sum": np.asarray([np.asarray(a, np.int64).sum() for a, _ in chunks], np.int64),
          "chunk_first": np.asarray([np.asarray(a)[0] for a, _ in chunks], np.int64),
          "chunk_last": np.asarray([np.asarray(a)[-1] for a, _ in chunks], np.int64),
          "chunk_len": np.asarray([len(a) for a, _ in chunks], np.int64)}
    rows = []
    for ci, (_, lab) in enumerate(chunks):
        for fr, evs in lab.items():
            for ev in evs:
                rows.append([ci, fr, ev[0], ev[1], ev[2], ev[3]])
    ck["chunk_label_rows"] = np.asarray(rows, np.float64)

    class _Stub:  # carries the attributes Dataset.sample_filelist_for_train_iter reads (datasets.py:67-92)
        pass
    names = [f"f{i:02d}" for i in range(23)]
    stub = _Stub()
    stub.total_filelist = list(names); stub.remaining_file = list(names); stub.nb_samples = 7; stub.filelist = []
    random.seed(5)
    epochs = []
    for _ in range(9):
        ref_datasets.Dataset.sample_filelist_for_train_iter(stub)
        epochs.append([names.index(x) for x in stub.filelist])
    ck["sampler_epochs"] = np.asarray(epochs, np.int64)
    np.savez_compressed(os.path.join(GOLD, "chunking.npz"), **ck)
    # ---- decode + conn-merge NMS (reference LabelPostProcessor on CPU)
    pn = ref_shims.ref_params(12)
    pn["train_config"].update({"conf_thresh": 0.5, "clss_thresh": 0.5, "unify_thresh": 15., "nms": "conn-merge"})
    post = ref_datasets.LabelPostProcessor(pn)
    gn = torch.Generator().manual_seed(3)
    Tn = 24
    lg = torch.randn(1, Tn, 2400, generator=gn) * 1.5 - 1.0
    yv = lg.reshape(1, Tn, 8, 4, 5, 15)
    yv[0, :, 2, 1, :, 0] += 5; yv[0, :, 2, 1, :, 3] += 5                       # a cluster of class 2 in cell (2,1)
    yv[0, ::2, 3, 1, :2, 0] += 5; yv[0, ::2, 3, 1, :2, 3] += 5                # neighbours in the next cell
    yv[0, :, 6, 2, 0, 0] += 5; yv[0, :, 6, 2, 0, 8] += 5                       # an isolated class-7 detection
    dn = post.postprocess(lg.clone())                                          # datasets.py:741-857
    rows = [[fr] + d for fr, dets in dn.items() for d in dets]
    extra = {}
    for mode in ("soft-merge", "nms"):                                         # datasets.py:817-846
        pn["train_config"]["nms"] = mode
        dm = ref_datasets.LabelPostProcessor(pn).postprocess(lg.clone())
        extra["rows_" + mode.replace("-", "_")] = np.asarray([[fr] + d for fr, dets in dm.items() for d in dets], np.float64)
    np.savez_compressed(os.path.join(GOLD, "nms_ref.npz"), logit=lg.numpy(), rows=np.asarray(rows, np.float64), **extra)
    print("golden fixtures written to", GOLD)
    for fn in sorted(os.listdir(GOLD)):
        print(" ", fn, os.path.getsize(os.path.join(GOLD, fn)))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "overlap":
        gen_overlap_cells()
    else:
        main()
        gen_overlap_cells()
