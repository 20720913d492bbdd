"""Data path either side of the hot path: on-the-fly chunk views, epoch
sampler and file formats against the unmodified reference (golden chunking.npz)."""
import random

import numpy as np
import pytest
import torch

from conftest import ROOT  # noqa: F401


@pytest.fixture(scope="module")
def A():
    import __graft_entry__ as g
    g.build()
    import adyolo_b200
    return adyolo_b200


def _label(frames, events):
    lab = {}
    for f, e in zip(frames, events):
        lab.setdefault(int(f), []).append([int(e[0]), int(e[1]), float(e[2]), float(e[3])])
    return lab


def _file(g):
    rng = np.random.default_rng(int(g["audio_seed"]))
    return rng.integers(-3000, 3000, size=(int(g["n_samples"]), 4)).astype(np.int16)


def test_chunk_views_match_reference_chunk_instance(A, gold):
    g = gold("chunking.npz")
    audio = _file(g)
    lab = _label(g["label_frames"], g["label_events"])
    store = A.ResidentClips(device="cpu")
    store.add("fold1_room1_mix001", audio, lab)
    store.finalize()
    names = store.chunk_names()
    assert len(names) == int(g["n_chunks"]) and names[0] == "fold1_room1_mix001_chunk001"
    offs, events, N, nlf = store.batch(names)
    assert N == int(g["chunk_len"][0]) and nlf == 200
    res = store.audio.numpy().astype(np.int64)
    for ci, o in enumerate(offs.tolist()):
        sl = res[o:o + N]
        assert sl.sum() == g["chunk_sum"][ci]
        assert (sl[0] == g["chunk_first"][ci]).all() and (sl[-1] == g["chunk_last"][ci]).all()
    want = g["chunk_label_rows"]                       # [chunk, frame', class, src, azi, ele]
    got = events.numpy()                               # [batch, frame', class, azi, ele]
    assert np.array_equal(got, want[:, [0, 1, 2, 4, 5]])


def test_epoch_sampler_reproduces_reference_sequence(A, gold):
    g = gold("chunking.npz")
    names = [f"f{i:02d}" for i in range(23)]
    smp = A.EpochSampler(names, 7)
    random.seed(5)
    for want in g["sampler_epochs"]:
        got = [names.index(x) for x in smp.sample_filelist_for_train_iter()]
        assert got == list(want)
    rem = smp.get_remaining_file()
    smp2 = A.EpochSampler(names, 7)
    smp2.init_remaining_file_from_list(list(rem))       # resume from a checkpointed pool
    assert smp2.get_remaining_file() == rem


def test_wav_and_csv_formats(A, tmp_path):
    import scipy.io.wavfile as wav
    rng = np.random.default_rng(0)
    a = rng.integers(-32768, 32767, size=(4800, 4)).astype(np.int16)
    wav.write(tmp_path / "x.wav", 24000, a)
    assert np.array_equal(A.load_wav2npy(tmp_path / "x.wav"), a)
    (tmp_path / "x.csv").write_text("0,3,0,-45,10\n0,5,1,180,-90\n7,3,0,12.5,0\n")
    lab = A.load_csv2dict(tmp_path / "x.csv")
    assert lab == {0: [[3, 0, -45.0, 10.0], [5, 1, 180.0, -90.0]], 7: [[3, 0, 12.5, 0.0]]}


@pytest.mark.gpu
def test_view_features_equal_materialised_chunks(A, gold, scaler2021):
    from adyolo_b200.features import _scaler_to_device
    g = gold("chunking.npz")
    audio = _file(g)
    store = A.ResidentClips(device="cuda")
    store.add("a", audio, _label(g["label_frames"], g["label_events"]))
    store.add("b", audio[::-1].copy(), {})
    store.finalize()
    names = [n for n in store.chunk_names() if n.endswith(("chunk001", "chunk004", "chunk008"))]
    offs, events, N, nlf = store.batch(names)
    sd = _scaler_to_device(scaler2021, ("MEL", "IV"), torch.device("cuda"))
    rot = torch.arange(len(names), dtype=torch.int8, device="cuda") % 16
    via_views = A.features_batched_views(store.audio, offs, N, sd, rot_comb=rot)
    dense = torch.stack([store.audio[o:o + N] for o in offs.tolist()])
    assert torch.equal(via_views, A.features_batched(dense, sd, rot_comb=rot))
    rows = A.label_rows_batched(events, nlf, A.labels.GridSpec(12, 5, [45, 45], 0.5), rot_comb=rot)
    assert rows.shape[1] == 7 and rows.shape[0] > 0


def test_cpulist_parser_for_numa_binding(A):
    from adyolo_b200.pipeline import _parse_cpulist
    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist("") == set()
    assert _parse_cpulist("5") == {5}
