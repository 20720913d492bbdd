"""CPU-side checks: the C-ABI library loads and exports every declared symbol (no compute calls
without a GPU), host helpers match the oracle, and the CPU emulation of the front-end kernel
(same per-thread code as the CUDA kernel, built with g++) matches the oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from oracle import features_np as F


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    import adyolo_b200
    return adyolo_b200


def test_library_exports_every_declared_symbol(built):
    from adyolo_b200 import _lib
    L = _lib.lib()
    syms = _lib.declared_symbols()
    assert len(syms) >= 15
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert L.adyolo_version() >= 100
    # every declared symbol has a ctypes signature on the Python side
    assert not [s for s in syms if s not in _lib._SIGS]


def test_sass_is_sm100a(built):
    from adyolo_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_mel_filterbank_host_is_bitexact(built):
    m = built.mel_filterbank(24000, 1200, 64)
    assert np.array_equal(m, F.librosa_mel(24000, 1200, 64))
    m2 = built.mel_filterbank(16000, 512, 40)
    assert np.abs(m2 - F.librosa_mel(16000, 512, 40)).max() < 1e-9


def test_unsupported_geometry_is_refused(built):
    from adyolo_b200 import _lib
    cfg = _lib.FrontendCfg(24000, 1024, 512, 1024, 64, 4, 1e-8, 80.0)
    assert _lib.lib().adyolo_frontend_workspace_bytes(ctypes.byref(cfg), 1, 24000) == 0
    assert b"1200" in _lib.lib().adyolo_last_error()


def test_grids_beyond_the_compiled_limits_are_refused(built):
    """one 32-bit cell mask per event (label path), 16 x 16 cells (loss).  A finer grid than the
    reference's 45 degrees is refused with ADYOLO_ERR_UNSUPPORTED before anything is launched (host-side check: runs
    without a GPU), never truncated."""
    from adyolo_b200 import _lib
    from adyolo_b200.labels import GridSpec
    L = _lib.lib()
    total = ctypes.c_int64(-1)
    g30 = GridSpec(12, 5, [30, 30], 0.5)                       # 12 x 6 = 72 cells
    assert g30.nb_grids == [12, 6]
    rc = L.adyolo_label_cells(None, 0, 50, ctypes.byref(g30.c), None, 0, None, ctypes.byref(total), None, None)
    assert rc == -3 and b"32-cell" in L.adyolo_last_error()
    rc = L.adyolo_label_cells_rows(None, 0, 50, ctypes.byref(g30.c), None, 0, None, ctypes.byref(total), None, None, 0, None)
    assert rc == -3 and b"32-cell" in L.adyolo_last_error()
    g20 = GridSpec(12, 5, [20, 20], 0.5)                       # 18 x 9: beyond the loss kernels' 16 x 16 too
    rc = L.adyolo_label_cells(None, 0, 50, ctypes.byref(g20.c), None, 0, None, ctypes.byref(total), None, None)
    assert rc == -3 and b"too large" in L.adyolo_last_error()
    assert L.adyolo_loss_workspace_bytes(1, 1, ctypes.byref(g20.c)) == 0
    assert total.value == -1                                   # nothing was written


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        built.features_batched(torch.zeros((1, 24000, 4), dtype=torch.int16))
    with pytest.raises(RuntimeError):
        built.audio2stft(np.zeros((2400, 4)), 4, 1200, 600, 1200)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ad-yolo_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "/oracle/" not in txt, fn


# ---------------------------------------------------------------- kernel logic on the CPU
def _emu():
    L = ctypes.CDLL(os.path.join(ROOT, "build", "emu_frontend.so"))
    L.emu_features_foa.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_void_p]
    return L


# Intensity gates.  BASELINE.json: "1e-3 absolute on intensity channels" -- met on the raw channel values with
# a 20x margin for every fixture (RAW_IV_TOL).  The stricter reading of SURVEY H5 (1e-3 on the STANDARDISED output,
# i.e. raw error / std with std 0.005..0.015) holds for 'noise' and 'bursts'; the synthetic 'harsh' clip (full-scale
# tones over a 2-LSB dither, 75 dB inside one frame) cannot meet it in float32 at all: rounding the windowed int16
# frame to float32 alone costs 5.0e-4 there and a float32 pocketfft costs 2.1e-3 whichever way channels or frames
# are paired (tools/fp32_floor.py -> profiles/r02_fp32_floor.txt; exception recorded in BASELINE.md section 5).
RAW_IV_TOL = 5e-5


@pytest.mark.parametrize("name,tol_iv", [("noise", 1e-3), ("bursts", 1e-3), ("harsh", 2.5e-3)])
def test_emulated_kernel_matches_golden(built, gold, scaler2021, name, tol_iv):
    g = gold("features_foa.npz")
    clip = np.ascontiguousarray(g[f"{name}_audio"])
    N = len(clip); T = N // 600
    mel = built.mel_filterbank(24000, 1200, 64).copy()
    mean = np.concatenate([scaler2021["MEL"]["mean"][0].T, scaler2021["IV"]["mean"][0].T], 0).astype(np.float32)
    istd = (1.0 / np.concatenate([scaler2021["MEL"]["std"][0].T, scaler2021["IV"]["std"][0].T], 0)).astype(np.float32)
    out = np.zeros((1, 7, T, 64), np.float32)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert _emu().emu_features_foa(vp(clip), 1, N, vp(mel), vp(mean), vp(istd), 1e-8, 80.0, vp(out)) == 0
    ref = np.concatenate([g[f"{name}_MEL"].transpose(2, 0, 1), g[f"{name}_IV"].transpose(2, 0, 1)], 0)
    assert (np.abs(out[0, :4] - ref[:4]) / np.maximum(np.abs(ref[:4]), 1.0)).max() < 1e-4
    assert np.abs(out[0, 4:] - ref[4:]).max() < tol_iv
    istd_iv = istd[4:, None, :]
    assert np.abs((out[0, 4:] - ref[4:]) / istd_iv).max() < RAW_IV_TOL


def test_emulated_kernel_ragged_batch(built):
    """T not a multiple of the 3-frame tile, several clips, raw (un-standardised) output."""
    rng = np.random.default_rng(5)
    B, N = 3, 600 * 7 + 123
    clips = np.clip(rng.standard_normal((B, N, 4)) * 2000, -32768, 32767).astype(np.int16)
    T = N // 600
    mel = built.mel_filterbank(24000, 1200, 64).copy()
    out = np.zeros((B, 7, T, 64), np.float32)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert _emu().emu_features_foa(vp(clips), B, N, vp(mel), None, None, 1e-8, 80.0, vp(out)) == 0
    for b in range(B):
        ref = F.features_foa_stack(clips[b])
        assert ref.shape == (7, T, 64)
        assert (np.abs(out[b, :4] - ref[:4]) / np.maximum(np.abs(ref[:4]), 1.0)).max() < 1e-4
        assert np.abs(out[b, 4:] - ref[4:]).max() < 1e-6


def test_codelets_against_numpy_fft(built):
    L = ctypes.CDLL(os.path.join(ROOT, "build", "emu_codelets.so"))
    rng = np.random.default_rng(0)
    x = rng.standard_normal(1200) + 1j * rng.standard_normal(1200)
    i = np.stack([x.real, x.imag], 1).astype(np.float32).copy()
    o = np.zeros_like(i)
    L.emu_fft1200(i.ctypes.data_as(ctypes.c_void_p), o.ctypes.data_as(ctypes.c_void_p))
    ref = np.fft.fft(x)
    assert np.abs((o[:, 0] + 1j * o[:, 1]) - ref).max() / np.abs(ref).max() < 1e-6


# ---------------------------------------------------------------- second-generation kernel (fe2) on the CPU
def _emu2():
    L = ctypes.CDLL(os.path.join(ROOT, "build", "emu_fe2.so"))
    L.emu_fe2_features_foa.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p]
    return L


@pytest.mark.parametrize("name,tol_iv", [("noise", 1e-3), ("bursts", 1e-3), ("harsh", 2.5e-3)])
def test_emulated_fe2_kernel_matches_golden(built, gold, scaler2021, name, tol_iv):
    """The per-thread code of the product kernel (fe2_core.cuh: staging map, 16 x 15 x 5 index maps, in-place V layout,
    balanced mel schedule) run thread by thread on the CPU against the reference-derived golden features."""
    g = gold("features_foa.npz")
    clip = np.ascontiguousarray(g[f"{name}_audio"])
    N = len(clip); T = N // 600
    mel = built.mel_filterbank(24000, 1200, 64).copy()
    mean = np.concatenate([scaler2021["MEL"]["mean"][0].T, scaler2021["IV"]["mean"][0].T], 0).astype(np.float32)
    istd = (1.0 / np.concatenate([scaler2021["MEL"]["std"][0].T, scaler2021["IV"]["std"][0].T], 0)).astype(np.float32)
    out = np.zeros((1, 7, T, 64), np.float32)
    cost = (ctypes.c_long * 2)()
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert _emu2().emu_fe2_features_foa(vp(clip), 1, N, vp(mel), vp(mean), vp(istd), 1e-8, 80.0, None, vp(out), cost) == 0
    ref = np.concatenate([g[f"{name}_MEL"].transpose(2, 0, 1), g[f"{name}_IV"].transpose(2, 0, 1)], 0)
    assert (np.abs(out[0, :4] - ref[:4]) / np.maximum(np.abs(ref[:4]), 1.0)).max() < 1e-4
    assert np.abs(out[0, 4:] - ref[4:]).max() < tol_iv
    assert np.abs((out[0, 4:] - ref[4:]) / istd[4:, None, :]).max() < RAW_IV_TOL
    assert cost[0] <= 2 * cost[1]          # mel gather: at most 2x the conflict-free wavefront count


def test_emulated_fe2_ragged_batch_and_rotation(built):
    """Odd frame count (last tile holds one frame), several clips, raw output, and the 16 rotation combinations
    (channel signs / swap applied inside the kernel) against the oracle on explicitly rotated audio."""
    from oracle import augment_np
    rng = np.random.default_rng(6)
    B, N = 16, 600 * 7 + 321
    clips = np.clip(rng.standard_normal((B, N, 4)) * 2000, -32767, 32767).astype(np.int16)
    clips[:, :900] = 0
    T = N // 600
    mel = built.mel_filterbank(24000, 1200, 64).copy()
    out = np.zeros((B, 7, T, 64), np.float32)
    bits = np.array([[0x0, 0x2, 0x1, 0x3, 0x5, 0x7, 0x4, 0x6, 0x9, 0xB, 0x8, 0xA, 0xC, 0xE, 0xD, 0xF][c] for c in range(16)], np.uint8)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert _emu2().emu_fe2_features_foa(vp(clips), B, N, vp(mel), None, None, 1e-8, 80.0, vp(bits), vp(out), None) == 0
    for b in range(B):
        rot, _ = augment_np.rotate(clips[b], {}, b)
        ref = F.features_foa_stack(rot)
        assert ref.shape == (7, T, 64)
        assert (np.abs(out[b, :4] - ref[:4]) / np.maximum(np.abs(ref[:4]), 1.0)).max() < 1e-4, b
        assert np.abs(out[b, 4:] - ref[4:]).max() < 1e-6, b


def test_emulated_fe2_mic_phasors_give_the_oracle_gcc(built):
    """MIC format: the unit phasors stage C hands to the GCC-PHAT kernel (position order, conjugate for folded bins,
    channel order) reproduce the float64 GCC-PHAT oracle when the lag transform is done in numpy."""
    L = ctypes.CDLL(os.path.join(ROOT, "build", "emu_fe2.so"))
    rng = np.random.default_rng(9)
    N = 600 * 9
    clip = np.clip(rng.standard_normal((N, 4)) * 1500, -32767, 32767).astype(np.int16)
    clip[:, 1] = np.roll(clip[:, 0], 3)                              # a delayed copy: a clear GCC peak at lag 3
    T = N // 600
    mel = built.mel_filterbank(24000, 1200, 64).copy()
    ph = np.zeros((T, 608, 4, 2), np.float32)
    bop = np.zeros(608, np.int32)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert L.emu_fe2_mic_phasors(vp(clip), ctypes.c_longlong(N), vp(mel), ctypes.c_float(1e-8), vp(ph), vp(bop)) == 0
    assert sorted(bop[bop >= 0].tolist()) == list(range(601))        # every bin exactly once
    u = np.zeros((T, 601, 4), np.complex128)
    for p in range(608):
        if bop[p] >= 0:
            u[:, bop[p]] = ph[:, p, :, 0] + 1j * ph[:, p, :, 1]
    assert np.abs(np.abs(u) - 1).max() < 1e-5
    got = []
    for m in range(4):
        for n in range(m + 1, 4):
            cc = np.fft.irfft(np.conj(u[:, :, m]) * u[:, :, n], n=1200, axis=1)
            got.append(np.concatenate((cc[:, -32:], cc[:, :32]), axis=-1))
    got = np.stack(got, -1)
    audio = F.normalise_int16(clip)
    ref = F.gcc_phat(F.audio2stft(audio, T, 1200, 600, 1200, "han"), 1200, 64)
    assert np.abs(got - ref).max() < 1e-4
    assert np.argmax(ref[2, :, 0]) == 32 + 3


def test_tile_coordinate_magic_multiply():
    """fe2.cu::tile_coords replaces `tile / tiles_per_clip` by __umulhi(tile, magic) + one correction, magic =
    min(floor(2^32 / tpc), 2^32 - 1) (host side of launch_inst).  The estimate may undershoot the quotient by at most one
    for every tile < 2^31: checked here on the arithmetic itself (edges of every quotient step for a spread of divisors)."""
    rng = np.random.default_rng(0)
    tpcs = [1, 2, 3, 5, 100, 125, 400, 1200, 65535, 65536, 65537, 720000, 2**20 + 1, 2**30, 2**31 - 1]
    tpcs += [int(x) for x in rng.integers(1, 2**31 - 1, size=40)]
    for tpc in tpcs:
        magic = min((1 << 32) // tpc, (1 << 32) - 1)
        tiles = {0, 1, tpc - 1, tpc, tpc + 1, 2**31 - 1, 2**31 - 2}
        for k in [int(x) for x in rng.integers(0, max((2**31 - 1) // tpc, 1), size=30)]:
            tiles.update((k * tpc - 1, k * tpc, k * tpc + tpc - 1))
        for tile in tiles:
            if not 0 <= tile < 2**31:
                continue
            q = (tile * magic) >> 32
            r = tile - q * tpc
            assert 0 <= r < 2 * tpc, (tpc, tile)
            if r >= tpc:
                q, r = q + 1, r - tpc
            assert (q, r) == divmod(tile, tpc), (tpc, tile)
