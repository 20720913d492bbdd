"""Pin the feature oracle: against the reference's own classes executed here (golden npz),
and the restated librosa arithmetic against independent torch / torchaudio implementations."""
import numpy as np
import pytest
import torch

from oracle import features_np as F


@pytest.mark.parametrize("name", ["noise", "bursts", "harsh"])
def test_oracle_matches_reference_golden(gold, scaler2021, name):
    g = gold("features_foa.npz")
    clip = g[f"{name}_audio"]
    (MEL, IV), nlf = F.features_foa(F.normalise_int16(clip), scaler=scaler2021)
    assert nlf == int(g[f"{name}_nlf"])
    np.testing.assert_array_equal(MEL, g[f"{name}_MEL"])
    np.testing.assert_array_equal(IV, g[f"{name}_IV"])
    (mel_raw, iv_raw), _ = F.features_foa(F.normalise_int16(clip))
    np.testing.assert_array_equal(mel_raw, g[f"{name}_mel_raw"])
    np.testing.assert_array_equal(iv_raw, g[f"{name}_iv_raw"])
    T = len(clip) // 600
    spec = F.audio2stft(F.normalise_int16(clip), T, 1200, 600, 1200)
    assert spec.shape == (T, 601, 4) and spec.dtype == np.complex128
    np.testing.assert_array_equal(spec[::17, ::13, :], g[f"{name}_spec_probe"])


def test_stft_against_torch_stft(gold):
    clip = gold("features_foa.npz")["bursts_audio"]
    a = F.normalise_int16(clip)
    T = len(a) // 600
    spec = F.audio2stft(a, T, 1200, 600, 1200)
    w = torch.hann_window(1200, periodic=True, dtype=torch.float64)
    for c in range(4):
        ts = torch.stft(torch.from_numpy(a[:, c]), 1200, 600, 1200, window=w, center=True,
                        pad_mode="reflect", return_complex=True).numpy()
        assert ts.shape[1] == T + 1                      # librosa: 1 + N//hop frames; last dropped
        assert np.abs(ts[:, :T].T - spec[:, :, c]).max() < 1e-12


def test_mel_against_torchaudio():
    ta = pytest.importorskip("torchaudio")
    fb = ta.functional.melscale_fbanks(601, 0.0, 12000.0, 64, 24000, norm="slaney", mel_scale="slaney").numpy()
    mel = F.librosa_mel(24000, 1200, 64)
    assert mel.dtype == np.float32 and mel.shape == (64, 601)
    assert np.abs(fb - mel.T).max() < 1e-7
    assert int((mel != 0).sum()) == 1165
    nz = (mel != 0)
    assert nz.sum(0).max() <= 2 and nz.sum(1).min() >= 5  # each bin feeds <=2 mels


def test_power_to_db_topdb_is_global_per_channel(gold):
    g = gold("features_foa.npz")
    raw = g["bursts_mel_raw"]
    for c in range(4):
        assert raw[:, :, c].min() >= raw[:, :, c].max() - 80.0 - 1e-9
    assert (raw == raw.max(axis=(0, 1), keepdims=True) - 80.0).any()  # clamp really active


def test_gcc_phat_properties():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4800, 4)) * 0.1
    x[:, 1] = np.roll(x[:, 0], 5)                       # channel 1 = channel 0 delayed by 5
    spec = F.audio2stft(x, 6, 1200, 600, 1200)
    g = F.gcc_phat(spec, 1200, 64)
    assert g.shape == (6, 64, 6)
    assert (np.argmax(g[2:5, :, 0], axis=1) == 32 + 5).all()  # pair (0,1): peak at lag +5
