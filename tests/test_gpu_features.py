"""GPU parity of the feature front end against the oracle / reference-generated goldens.
All calls go through the C ABI (adyolo_b200 Python mirror -> ctypes -> libadyolo_b200.so).

Tolerances (BASELINE.json north_star, made well-posed as SURVEY H5 recommends):
  log-mel : |a-b| <= 1e-4 * max(|b|, 1)   on dB values, before and after standardisation
  IV      : |a-b| <= 1e-3 absolute         on the standardised output (raw IV is <= 0.044)
The 'harsh' clip (channels 75 dB apart inside a frame) is checked at IV 2.5e-3: the FP32 pipeline's
dynamic-range floor (measured in tools/fp32_floor.py, stated in BASELINE.md section 5 and DESIGN.md section 2)."""
import numpy as np
import pytest
import torch

from oracle import features_np as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    import __graft_entry__ as g
    g.build()
    import adyolo_b200
    assert torch.cuda.is_available()
    return adyolo_b200


def _scaler_dev(A, scaler):
    from adyolo_b200.features import _scaler_to_device
    return _scaler_to_device(scaler, ("MEL", "IV"), torch.device("cuda"))


def _mel_err(a, b):
    return (np.abs(a - b) / np.maximum(np.abs(b), 1.0)).max()


# Intensity gates.  BASELINE.json: "1e-3 absolute on intensity channels" -- met on the raw channel values with
# a 20x margin for every fixture (RAW_IV_TOL).  The stricter reading of SURVEY H5 (1e-3 on the STANDARDISED output,
# i.e. raw error / std with std 0.005..0.015) holds for 'noise' and 'bursts'; the synthetic 'harsh' clip (full-scale
# tones over a 2-LSB dither, 75 dB inside one frame) cannot meet it in float32 at all: rounding the windowed int16
# frame to float32 alone costs 5.0e-4 there and a float32 pocketfft costs 2.1e-3 whichever way channels or frames
# are paired (tools/fp32_floor.py -> profiles/r02_fp32_floor.txt; exception recorded in BASELINE.md section 5).
RAW_IV_TOL = 5e-5


@pytest.mark.parametrize("name,tol_iv", [("noise", 1e-3), ("bursts", 1e-3), ("harsh", 2.5e-3)])
def test_fused_features_vs_reference_golden(A, gold, scaler2021, name, tol_iv):
    g = gold("features_foa.npz")
    clip = torch.from_numpy(g[f"{name}_audio"]).cuda()
    raw = A.features_batched(clip[None], None).cpu().numpy()[0]
    ref_raw = np.concatenate([g[f"{name}_mel_raw"].transpose(2, 0, 1), g[f"{name}_iv_raw"].transpose(2, 0, 1)], 0)
    assert raw.shape == ref_raw.shape
    assert _mel_err(raw[:4], ref_raw[:4]) < 1e-4
    assert np.abs(raw[4:] - ref_raw[4:]).max() < RAW_IV_TOL         # BASELINE.json's gate is 1e-3 on these values
    std = A.features_batched(clip[None], _scaler_dev(A, scaler2021)).cpu().numpy()[0]
    ref = np.concatenate([g[f"{name}_MEL"].transpose(2, 0, 1), g[f"{name}_IV"].transpose(2, 0, 1)], 0)
    assert _mel_err(std[:4], ref[:4]) < 1e-4
    assert np.abs(std[4:] - ref[4:]).max() < tol_iv


def test_topdb_clamp_is_exercised_and_exact(A, gold):
    g = gold("features_foa.npz")
    clip = torch.from_numpy(g["bursts_audio"]).cuda()
    raw = A.features_batched(clip[None], None).cpu().numpy()[0]
    un = A.features_batched(clip[None], None, apply_topdb=False).cpu().numpy()[0]
    for c in range(4):
        mx = un[c].max()
        assert np.array_equal(raw[c], np.maximum(un[c], np.float32(mx) - np.float32(80.0)))
        assert (un[c] < mx - 80).any()            # the clamp really bites on this clip
    assert np.array_equal(raw[4:], un[4:])


def test_ragged_batch_and_tile_tail(A):
    rng = np.random.default_rng(11)
    for N in (600 * 7 + 123, 601, 1200, 600 * 3, 600 * 4 + 1):
        B = 3
        clips = np.clip(rng.standard_normal((B, N, 4)) * 2500, -32768, 32767).astype(np.int16)
        out = A.features_batched(torch.from_numpy(clips).cuda(), None).cpu().numpy()
        assert out.shape == (B, 7, N // 600, 64)
        for b in range(B):
            ref = F.features_foa_stack(clips[b])
            assert _mel_err(out[b, :4], ref[:4]) < 1e-4
            assert np.abs(out[b, 4:] - ref[4:]).max() < 5e-6


def test_digital_silence_and_full_scale(A):
    """amin / 1e-8 DC offset path (all-zero clip) and int16 extremes.  The full-scale Nyquist
    square wave puts everything in bin 600 and leaves the other 63 mel bands ~75 dB down on
    leakage + a -0.5 LSB DC term: that is the FP32 dynamic-range floor again, hence
    the looser 5e-4 there."""
    N = 24000
    z = np.zeros((N, 4), np.int16)
    fs = np.full((N, 4), -32768, np.int16); fs[::2] = 32767
    for clip, tol, tol_iv in ((z, 1e-4, 5e-6), (fs, 5e-4, 2e-3)):
        out = A.features_batched(torch.from_numpy(clip).cuda()[None], None).cpu().numpy()[0]
        ref = F.features_foa_stack(clip)
        assert np.isfinite(out).all()
        assert _mel_err(out[:4], ref[:4]) < tol
        # full-scale Nyquist: the exact spectrum is ZERO in bins 2..598, so the reference's I/E is
        # 0/eps there while FP32 leaves ~1e-7 * full-scale leakage that eps=1e-8 does not mask
        assert np.abs(out[4:] - ref[4:]).max() < tol_iv


def test_per_clip_reference_surface(A, gold, scaler2021, tmp_path):
    """FeatureLabelProcessor.get_feature / audio2stft / stft2melscale / stft2iv (reference signatures)."""
    import pickle
    from oracle.loss_torch import default_params
    g = gold("features_foa.npz")
    with open(tmp_path / "scaler_wts.pkl", "wb") as f:
        pickle.dump(scaler2021, f)
    p = default_params(12, "cuda:0")
    p["data_config"]["data_pth"] = str(tmp_path)
    flp = A.FeatureLabelProcessor(p)
    clip = g["bursts_audio"]
    audio = clip / 32768.0 + 1e-8
    (MEL, IV), nlf = flp.get_feature(audio)
    assert nlf == int(g["bursts_nlf"]) and MEL.shape == g["bursts_MEL"].shape and MEL.dtype == np.float64
    assert _mel_err(MEL, g["bursts_MEL"]) < 1e-4
    assert np.abs(IV - g["bursts_IV"]).max() < 1e-3
    T = len(audio) // 600
    spec = A.audio2stft(audio, T, 1200, 600, 1200, "han")
    assert spec.shape == (T, 601, 4)
    ref_spec = F.audio2stft(audio, T, 1200, 600, 1200)
    assert np.abs(spec - ref_spec).max() < 2e-6 * np.abs(ref_spec).max()
    np.testing.assert_allclose(spec[::17, ::13, :], g["bursts_spec_probe"], atol=2e-6 * np.abs(ref_spec).max())
    mel = A.stft2melscale(ref_spec, 24000, 1200, 64)
    assert _mel_err(mel, g["bursts_mel_raw"]) < 1e-4
    iv = A.stft2iv(ref_spec, 24000, 1200, 64)
    assert np.abs(iv - g["bursts_iv_raw"]).max() < 5e-6
    # float (non-int16-representable) input goes through the float32 kernels
    a2 = audio * 0.731
    (M2, I2), _ = flp.get_feature(a2)
    (Mr, Ir), _ = F.features_foa(a2, scaler=scaler2021)
    assert _mel_err(M2, Mr) < 1e-4 and np.abs(I2 - Ir).max() < 1e-3


def test_error_behaviour(A):
    with pytest.raises(NotImplementedError):
        A.audio2stft(np.zeros((4800, 4)), 8, 1024, 512, 1024)
    with pytest.raises(NotImplementedError):
        A.audio2stft(np.zeros((4800, 4)), 8, 1200, 600, 1200, window="hamming")
    with pytest.raises(ValueError):
        A.features_batched(torch.zeros((2, 4800, 3), dtype=torch.int16, device="cuda"))
    with pytest.raises(RuntimeError):
        A.features_batched(torch.zeros((1, 600, 4), dtype=torch.int16, device="cuda"))   # reflect pad needs N > 600


def test_full_size_batch_properties(A, scaler2021):
    """BASELINE config 2 shape (256 x 5 s): size-independent properties instead of the slow oracle:
    batch independence (a clip's features do not depend on its batch neighbours or position),
    determinism, the top_db floor, and oracle parity on a few sampled clips."""
    g = torch.Generator(device="cuda").manual_seed(3)
    audio = (torch.randn((256, 120000, 4), device="cuda", generator=g) * 3000).clamp_(-32768, 32767).to(torch.int16)
    audio[5, 30000:60000] = 0
    sd = _scaler_dev(A, scaler2021)
    out = A.features_batched(audio, sd)
    assert out.shape == (256, 7, 200, 64) and torch.isfinite(out).all()
    assert torch.equal(out, A.features_batched(audio, sd))
    perm = torch.randperm(256, device="cuda", generator=g)
    assert torch.equal(A.features_batched(audio[perm], sd), out[perm])
    assert torch.equal(A.features_batched(audio[7:8], sd)[0], out[7])
    raw = A.features_batched(audio, None)
    mx = raw[:, :4].amax(dim=(2, 3), keepdim=True)
    assert (raw[:, :4] >= mx - 80.0).all()
    assert (raw[5, :4] == mx[5] - 80.0).any()
    for b in (0, 5, 255):
        ref = F.features_foa_stack(audio[b].cpu().numpy(), scaler=scaler2021)
        o = out[b].cpu().numpy()
        assert _mel_err(o[:4], ref[:4]) < 1e-4 and np.abs(o[4:] - ref[4:]).max() < 1e-3


def test_mic_gcc_path_selfconsistent(A):
    """MIC log-mel + GCC-PHAT: no reference implementation exists -> parity unpinned;
    checked against the numpy restatement of the upstream DCASE-baseline semantics."""
    from adyolo_b200.features import features_mic_batched
    rng = np.random.default_rng(2)
    B, N = 2, 24000
    x = rng.standard_normal((B, N, 4)) * 2000
    x[:, :, 1] = np.roll(x[:, :, 0], 7, axis=1) + rng.standard_normal((B, N)) * 50
    clips = np.clip(x, -32768, 32767).astype(np.int16)
    out = features_mic_batched(torch.from_numpy(clips).cuda()).cpu().numpy()
    assert out.shape == (B, 10, 40, 64)
    for b in range(B):
        ref = F.features_mic_stack(clips[b])
        assert _mel_err(out[b, :4], ref[:4]) < 1e-4
        assert np.abs(out[b, 4:] - ref[4:]).max() < 1e-3
        assert (out[b, 4, 5:35].argmax(-1) == 32 + 7).all()


def test_mic_product_path_silence_and_vanishing_channel(A):
    """fe2<MIC> -> gcc_ph16 (FP16 lag transform) on the cases the phasor hand-off treats specially: a vanishing channel
    spectrum is sent as the zero phasor (+0, +0) and every pair it takes part in becomes (1, 0), like upstream's
    np.angle(0) = 0.  Digital silence: all four channels carry only the 1e-8 DC offset (bins 0, +-1; every other bin is
    exactly zero in FP32), so every cross spectrum is (1, 0) and cc = delta at lag 0 -- exactly 1 with FP16 operands
    (the sum of 600 ones and two halves is exact).  The float64 oracle has rounding noise instead of exact zeros in
    those bins, so this case is a property test, not a parity test.  A clip that is silent for its first half checks
    that the zero handling is per (frame, position) and leaves the live frames within the usual gate."""
    from adyolo_b200.features import features_mic_batched
    rng = np.random.default_rng(5)
    N = 24000
    z = np.zeros((N, 4), np.int16)
    half = np.clip(rng.standard_normal((N, 4)) * 2000, -32768, 32767).astype(np.int16)
    half[:, 1] = np.roll(half[:, 0], 5)
    half[: N // 2] = 0
    out = features_mic_batched(torch.from_numpy(np.stack([z, half])).cuda()).cpu().numpy()
    assert out.shape == (2, 10, 40, 64) and np.isfinite(out).all()
    assert (out[0, 4:, :, 32] == 1.0).all()                                  # cc[0] of every pair and frame
    other = np.delete(out[0, 4:], 32, axis=-1)
    assert np.abs(other).max() < 1e-4                                        # FP16-rounded twiddles cancel to ~1e-5
    assert (out[1, 4:, :18, 32] == 1.0).all()                                # the silent half of the second clip
    ref = F.features_mic_stack(half)
    assert np.abs(out[1, 4:, 22:] - ref[4:, 22:]).max() < 1e-3               # live frames (frame 20 +- 1 straddles the edge)
    assert (out[1, 4, 22:38].argmax(-1) == 32 + 5).all()


def test_gcc_tensor_core_kernel_against_oracle(A):
    """adyolo_gcc_from_stft (tcgen05 TF32 lag transform) straight through the C ABI on synthetic
    spectra: random phases, a dead microphone, digital silence, magnitudes far outside the range
    where |x|^2 fits FP32, a frame count that leaves the last 64-frame tile ragged, and a
    channels-last output (generic-stride epilogue).  Gate 1e-3 absolute (north_star); observed 3e-5."""
    import ctypes as C
    from adyolo_b200 import _lib
    from adyolo_b200.features import _cfg, ptr, stream_ptr
    rng = np.random.default_rng(21)
    B, T = 2, 83                                             # 166 frames = 2 full tiles + 38
    spec = (rng.standard_normal((B, T, 601, 4)) + 1j * rng.standard_normal((B, T, 601, 4))).astype(np.complex64)
    spec[0, 3:6, :, 2] = 0                                   # dead microphone: its pairs -> delta at lag 0
    spec[0, 10:12] = 0                                       # silence
    spec[1, 5] *= 1e-25                                      # |x|^2 underflows FP32
    spec[1, 6] *= 1e+24                                      # |x|^2 overflows FP32
    spec[1, 7, 100:200] = 0
    ref = np.stack([F.gcc_phat(spec[b].astype(np.complex128), 1200, 64) for b in range(B)])   # (B, T, 64, 6)
    L, cfg = _lib.lib(), _cfg()
    d_spec = torch.from_numpy(spec).cuda()
    out = torch.empty(B, 6, T, 64, device="cuda")
    st = (C.c_int64 * 4)(6 * T * 64, T * 64, 64, 1)
    assert L.adyolo_gcc_from_stft(ptr(d_spec), B, T, C.byref(cfg), None, None, ptr(out), st, stream_ptr()) == 0
    got = out.permute(0, 2, 3, 1).cpu().numpy()
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() < 1e-3
    assert np.abs(got - ref).max() < 1e-4                    # tighter than the gate: TF32 operands are rounded, not truncated
    assert (got[0, 10:12, 32, :] == 1.0).all()               # silence: cc[0] exactly 1
    assert np.abs(got[0, 10:12, 33:, :]).max() < 5e-5        # TF32-rounded twiddles: sum_k cos(2 pi k l / N) cancels to ~1e-5
    # channels-last output with standardisation: generic-stride path == natural path
    mean = torch.from_numpy(rng.standard_normal((6, 64)).astype(np.float32) * 0.01).cuda()
    istd = torch.from_numpy((5 + rng.random((6, 64))).astype(np.float32)).cuda()
    nat = torch.empty(B, 6, T, 64, device="cuda")
    cl = torch.empty(B, T, 64, 6, device="cuda")
    st2 = (C.c_int64 * 4)(T * 64 * 6, 1, 64 * 6, 6)
    assert L.adyolo_gcc_from_stft(ptr(d_spec), B, T, C.byref(cfg), ptr(mean), ptr(istd), ptr(nat), st, stream_ptr()) == 0
    assert L.adyolo_gcc_from_stft(ptr(d_spec), B, T, C.byref(cfg), ptr(mean), ptr(istd), ptr(cl), st2, stream_ptr()) == 0
    assert torch.equal(cl.permute(0, 3, 1, 2), nat)
    want = (ref.transpose(0, 3, 1, 2) - mean.cpu().numpy()[None, :, None, :]) * istd.cpu().numpy()[None, :, None, :]
    assert np.abs(nat.cpu().numpy() - want).max() < 1e-3


def test_scaler_action_single_gpu(A):
    rng = np.random.default_rng(4)
    clips = [np.clip(rng.standard_normal((24000 * 2, 4)) * (500 + 800 * i), -32768, 32767).astype(np.int16) for i in range(5)]
    clips[2][10000:30000] = 0
    got = A.preprocess_scaler(clips, batch_clips=2)
    ref = F.scaler_stats(clips)
    for grp in ("MEL", "IV"):
        scale = ref[grp]["std"]
        assert got[grp]["mean"].shape == (1, 64, 4 if grp == "MEL" else 3)
        assert (np.abs(got[grp]["mean"] - ref[grp]["mean"]) <= 1e-5 * scale + 1e-9).all()
        assert (np.abs(got[grp]["std"] - ref[grp]["std"]) <= 1e-5 * scale + 1e-9).all()
        assert (np.abs(got[grp]["max"] - ref[grp]["max"]) <= 1e-4 * np.maximum(np.abs(ref[grp]["max"]), scale)).all()
        assert (np.abs(got[grp]["min"] - ref[grp]["min"]) <= 1e-4 * np.maximum(np.abs(ref[grp]["min"]), scale)).all()


def test_scaler_action_rel_1e6_on_eight_minutes(A):
    """statistics vs float64 numpy on a longer subset, relative 1e-6.  Mean / std are
    compared relative to max(|value|, std) of that (mel, channel) -- the intensity means are ~0, so a plain
    relative error is ill-posed there."""
    rng = np.random.default_rng(40)
    clips = []
    for i in range(8):                                   # 8 x 60 s, levels from -50 to -10 dBFS, one with a silent stretch
        x = rng.standard_normal((24000 * 60, 4)) * (100 * 2.0 ** i)
        if i == 3:
            x[200000:500000] = 0
        clips.append(np.clip(np.round(x), -32768, 32767).astype(np.int16))
    got = A.preprocess_scaler(clips, batch_clips=4)
    ref = F.scaler_stats(clips)
    worst = {}
    for grp in ("MEL", "IV"):
        scale = np.maximum(np.abs(ref[grp]["mean"]), ref[grp]["std"])
        worst[grp + ".mean"] = (np.abs(got[grp]["mean"] - ref[grp]["mean"]) / scale).max()
        worst[grp + ".std"] = (np.abs(got[grp]["std"] - ref[grp]["std"]) / scale).max()
    print("scaler rel errors:", worst)
    assert max(worst.values()) <= 1e-6, worst


def test_fused_rotation_augmentation_all_16_combinations(A, scaler2021):
    """RotationAug fused into the front end (channel signs / swap) and into the
    label kernel.  Oracle = reference rotation of the int16 audio + label dict (pinned numpy
    restatement) followed by the plain feature / cell oracles."""
    from oracle import augment_np, assign_np
    rng = np.random.default_rng(12)
    B, N = 16, 24000
    t = np.arange(N) / 24000.0
    s = 4000 * np.sin(2 * np.pi * 700 * t) * (t > 0.3)
    x = rng.standard_normal((B, N, 4)) * 300 + s[None, :, None] * np.array([1.0, 0.6, -0.3, 0.7])[None, None, :]
    x[:, :2400] = 0                                                # leading digital silence: DC term matters
    clips = np.clip(np.round(x), -32767, 32767).astype(np.int16)   # no -32768: the reference wraps it (int16)
    comb = np.arange(16, dtype=np.int8)
    sd = _scaler_dev(A, scaler2021)
    out = A.features_batched(torch.from_numpy(clips).cuda(), sd, rot_comb=torch.from_numpy(comb).cuda()).cpu().numpy()
    for b in range(B):
        a2, _ = augment_np.rotate(clips[b], {}, int(comb[b]))
        ref = F.features_foa_stack(np.ascontiguousarray(a2), scaler=scaler2021)
        assert _mel_err(out[b, :4], ref[:4]) < 1e-4, b
        assert np.abs(out[b, 4:] - ref[4:]).max() < 1e-3, b
    # labels
    grid = A.labels.GridSpec(12, 5, [45, 45], 0.5)
    E = 4000
    ev = np.stack([rng.integers(0, B, E), rng.integers(0, 12, E), rng.integers(0, 12, E),
                   rng.integers(-180, 181, E), rng.integers(-90, 91, E)], 1).astype(np.float64)
    rows = A.label_rows_batched(torch.from_numpy(ev).cuda(), 10, grid, rot_comb=torch.from_numpy(comb).cuda()).cpu().numpy()
    want = assign_np.events_to_rows(augment_np.rotate_events(ev, comb), 10).astype(np.float32)
    assert np.array_equal(rows, want)


def test_specaug_masks_on_device(A, scaler2021):
    """SURVEY §8(f) N2 (SpecAug half): the mask kernel applied to front-end output equals the oracle
    restatement of the reference's masking (pinned by tests/golden/specaug.npz on the CPU side);
    per-clip reference surface; MIC grouping; intervals clipped to the tensor."""
    import random
    from oracle import augment_np
    params = {"aug_config": {"spec_augment": True, "spec_augment_thresh": 0.7,
                             "spec_augment_time_mask_param": 40, "spec_augment_freq_mask_param": 40}}
    sa = A.SpecAug(params, is_valid=False)
    rng = np.random.default_rng(8)
    audio = torch.from_numpy(rng.integers(-3000, 3000, size=(6, 24000 * 2, 4)).astype(np.int16)).cuda()
    sd = _scaler_dev(A, scaler2021)
    feat = A.features_batched(audio, sd)
    ref_in = feat.cpu().numpy()
    random.seed(3); torch.manual_seed(3)
    rects = sa.draw(6, feat.shape[2], 64)
    assert rects.any()
    out = sa.apply_rects(feat.clone(), rects)
    np.testing.assert_array_equal(out.cpu().numpy(), augment_np.specaug_apply(ref_in, rects.numpy()))
    # batched draw+apply consumes the RNG exactly like draw()
    random.seed(3); torch.manual_seed(3)
    assert torch.equal(sa.augment_batched(feat.clone()), out)
    # per-clip reference surface: (C, T, F) group -> masked copy, input untouched
    random.seed(5); torch.manual_seed(5)
    r1 = sa.draw(1, feat.shape[2], 64, 1)
    random.seed(5); torch.manual_seed(5)
    grp = feat[2, :4].clone()
    y = sa.augment(grp)
    want = augment_np.specaug_apply(grp.cpu().numpy()[None], r1.numpy(), groups=((0, 4),))[0]
    np.testing.assert_array_equal(y.cpu().numpy(), want)
    assert torch.equal(grp, feat[2, :4])
    # MIC grouping (4 + 6 channels), odd F (no float4 path), out-of-range intervals are clipped
    x = torch.randn(2, 10, 37, 63, device="cuda")
    rr = torch.tensor([[[5, 20, 30, 99], [0, 0, -3, 2]], [[60, 70, 0, 0], [1, 2, 36, 37]]], dtype=torch.int32)
    got = sa.apply_rects(x.clone(), rr, groups=((0, 4), (4, 10))).cpu().numpy()
    np.testing.assert_array_equal(got, augment_np.specaug_apply(x.cpu().numpy(), np.clip(rr.numpy(), 0, None), groups=((0, 4), (4, 10))))
    assert A.SpecAug(params, is_valid=True).augment_batched(feat) is feat


def test_long_clips_batch_equals_per_clip_and_oracle(A, scaler2021):
    """60-s clips (BASELINE config[0] length) in a batch: every clip equals its single-clip result
    bit for bit (no cross-clip leakage; 2400 frames per clip = 800 front-end tiles, 37.5 GCC tiles so
    GCC tiles straddle clip boundaries), and one clip is checked against the float64 oracle."""
    from adyolo_b200.features import features_mic_batched
    rng = np.random.default_rng(33)
    B, N = 3, 24000 * 60
    x = rng.standard_normal((B, N, 4)) * 1500
    x[:, :, 1:] += x[:, :, :1] * 0.5
    x[1, 300000:340000] = 0
    audio = torch.from_numpy(np.clip(x, -32768, 32767).astype(np.int16)).cuda()
    sd = _scaler_dev(A, scaler2021)
    out = A.features_batched(audio, sd)
    assert out.shape == (B, 7, 2400, 64)
    for b in range(B):
        assert torch.equal(A.features_batched(audio[b:b + 1], sd)[0], out[b])
    ref = F.features_foa_stack(audio[1].cpu().numpy(), scaler=scaler2021)
    o = out[1].cpu().numpy()
    assert _mel_err(o[:4], ref[:4]) < 1e-4 and np.abs(o[4:] - ref[4:]).max() < 1e-3
    mic = features_mic_batched(audio)
    assert mic.shape == (B, 10, 2400, 64) and bool(torch.isfinite(mic).all())
    for b in range(B):
        assert torch.equal(features_mic_batched(audio[b:b + 1])[0], mic[b])
    refm = F.features_mic_stack(audio[2].cpu().numpy())
    m = mic[2].cpu().numpy()
    assert _mel_err(m[:4], refm[:4]) < 1e-4 and np.abs(m[4:] - refm[4:]).max() < 1e-3


def test_mic_standardisation_with_a_10_channel_scaler(A):
    """MIC path with a (10, 64) scaler (4 log-mel + 6 GCC rows, the layout the new `scaler` action writes
    for the MIC format): equals standardising the raw output (the top_db clamp commutes with it)."""
    from adyolo_b200.features import features_mic_batched
    rng = np.random.default_rng(44)
    clips = torch.from_numpy(np.clip(rng.standard_normal((3, 24000 * 2, 4)) * 2500, -32768, 32767).astype(np.int16)).cuda()
    mean = torch.from_numpy(np.concatenate([rng.uniform(-60, -20, (4, 64)), rng.normal(0, 0.01, (6, 64))]).astype(np.float32)).cuda()
    istd = torch.from_numpy(np.concatenate([rng.uniform(0.05, 0.2, (4, 64)), rng.uniform(5, 30, (6, 64))]).astype(np.float32)).cuda()
    raw = features_mic_batched(clips)
    std = features_mic_batched(clips, (mean, istd))
    want = (raw - mean[None, :, None, :]) * istd[None, :, None, :]
    err = (std - want).abs().amax(dim=(0, 2, 3)) / want.abs().amax(dim=(0, 2, 3))
    assert float(err.max()) < 1e-5


def test_scaler_action_mic_format(A):
    """`scaler` action for the MIC format: the pickle schema gains a GCC key (1, 64, 6);
    statistics equal float64 numpy statistics of the per-file features."""
    from adyolo_b200.features import features_mic_batched
    rng = np.random.default_rng(6)
    clips = [np.clip(rng.standard_normal((24000 * 2, 4)) * (800 + 500 * i), -32768, 32767).astype(np.int16) for i in range(4)]
    got = A.preprocess_scaler(clips, fmt="mic", batch_clips=2)
    assert set(got) == {"MEL", "GCC"} and got["GCC"]["mean"].shape == (1, 64, 6) and got["MEL"]["std"].shape == (1, 64, 4)
    feats = np.concatenate([features_mic_batched(torch.from_numpy(c[None]).cuda())[0].double().cpu().numpy() for c in clips], axis=1)  # (10, sumT, 64)
    for key, sl in (("MEL", slice(0, 4)), ("GCC", slice(4, 10))):
        x = feats[sl]                                            # (C, frames, 64)
        mean, std = x.mean(axis=1).T[None], x.std(axis=1).T[None]   # (1, 64, C)
        assert np.abs(got[key]["mean"] - mean).max() <= 1e-6 * max(1.0, np.abs(mean).max())
        assert np.abs(got[key]["std"] - std).max() <= 1e-6 * max(1.0, np.abs(std).max())
        assert np.array_equal(got[key]["max"], x.max(axis=1).T[None]) and np.array_equal(got[key]["min"], x.min(axis=1).T[None])


def test_nan_intensity_raises_instead_of_exit(A):
    """utility.py:211-213 prints and calls exit() when the intensity features contain NaN; the
    replacement reports it through the kernel's flag word and raises FloatingPointError."""
    rng = np.random.default_rng(1)
    spec = (rng.standard_normal((6, 601, 4)) + 1j * rng.standard_normal((6, 601, 4))).astype(np.complex128)
    assert np.isfinite(A.stft2iv(spec, 24000, 1200, 64)).all()
    spec[3, 100, 2] = np.nan
    with pytest.raises(FloatingPointError):
        A.stft2iv(spec, 24000, 1200, 64)
