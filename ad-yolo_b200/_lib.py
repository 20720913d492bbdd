"""Build + ctypes binding of libadyolo_b200.so (the C ABI in include/adyolo_b200.h)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libadyolo_b200.so")
SOURCES = ["fe2.cu", "frontend.cu", "frontend_aux.cu", "assign.cu", "labels.cu", "nms.cu", "gcc_tc.cu", "augment.cu", "tables.cu", "scaler.cu", "api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC"]


# build knobs: frames per tile of the round-1 kernel (csrc/frontend_core.cuh: 3 or 7), threads per CTA of the fe2
# kernel (csrc/fe2_core.cuh: 160 or 192)
if os.environ.get("ADY_TILE_FRAMES"):
    NVCC_FLAGS = NVCC_FLAGS + ["-DADY_TILE_FRAMES=" + os.environ["ADY_TILE_FRAMES"]]
if os.environ.get("ADY_FE2_NT"):
    NVCC_FLAGS = NVCC_FLAGS + ["-DADY_FE2_NT=" + os.environ["ADY_FE2_NT"]]
# experiments: load a variant build instead of the in-tree default (tools/build_variant.py)
if os.environ.get("ADYOLO_LIB"):
    LIB_PATH = os.environ["ADYOLO_LIB"]


def _nvcc():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link the shared library in-tree."""
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(ROOT, "build", "obj")
    os.makedirs(obj_dir, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "adyolo_b200.h")]
    newest = max(os.path.getmtime(d) for d in deps)
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= newest:
        return LIB_PATH

    def cc(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(cc, srcs))
    r = subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH, *objs],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


class FrontendCfg(C.Structure):
    _fields_ = [("sr", C.c_int32), ("n_fft", C.c_int32), ("hop_length", C.c_int32), ("win_length", C.c_int32),
                ("mel_bins", C.c_int32), ("n_channels", C.c_int32), ("dc_offset", C.c_float), ("top_db", C.c_float)]


class GridCfg(C.Structure):
    _fields_ = [("nb_classes", C.c_int32), ("nb_anchors", C.c_int32), ("grid_size", C.c_double * 2),
                ("g_overlap", C.c_double), ("n_thr", C.c_int32), ("train_unify", C.c_float * 4),
                ("angular_gain", C.c_float), ("object_gain", C.c_float), ("nonobj_gain", C.c_float),
                ("class_gain", C.c_float)]


_P = C.c_void_p
_SIGS = {
    "adyolo_last_error": (C.c_char_p, []),
    "adyolo_version": (C.c_int, []),
    "adyolo_mel_filterbank": (C.c_int, [C.c_int, C.c_int, C.c_int, _P]),
    "adyolo_frontend_workspace_bytes": (C.c_size_t, [C.POINTER(FrontendCfg), C.c_int, C.c_int64]),
    "adyolo_features_foa": (C.c_int, [_P, C.c_int, C.c_int64, C.POINTER(FrontendCfg), _P, _P, _P, _P, C.c_int, _P]),
    "adyolo_features_foa_rot": (C.c_int, [_P, C.c_int, C.c_int64, C.POINTER(FrontendCfg), _P, _P, _P, _P, _P, C.c_int, _P]),
    "adyolo_features_foa_views": (C.c_int, [_P, _P, C.c_int, C.c_int64, C.POINTER(FrontendCfg), _P, _P, _P, _P, _P, C.c_int, _P]),
    "adyolo_features_mic_logmel": (C.c_int, [_P, C.c_int, C.c_int64, C.POINTER(FrontendCfg), _P, _P, _P, _P, _P, C.c_int, _P]),
    "adyolo_features_mic_gcc": (C.c_int, [_P, C.c_int, C.c_int64, C.POINTER(FrontendCfg), _P, _P, _P, _P, _P, C.c_int, _P]),
    "adyolo_mic_spec_bytes": (C.c_size_t, [C.c_int, C.c_int64]),
    "adyolo_features_foa_clamp": (C.c_int, [_P, C.c_int, C.c_int64, C.POINTER(FrontendCfg), _P, _P, _P, _P]),
    "adyolo_stft": (C.c_int, [_P, C.c_int, C.c_int, C.c_int64, C.POINTER(FrontendCfg), _P, _P]),
    "adyolo_logmel_from_stft": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int, C.c_int, C.POINTER(FrontendCfg), _P, _P, _P,
                                          C.POINTER(C.c_int64), _P, C.c_int, _P]),
    "adyolo_iv_from_stft": (C.c_int, [_P, C.c_int, C.c_int64, C.POINTER(FrontendCfg), _P, _P, _P,
                                      C.POINTER(C.c_int64), _P, _P]),
    "adyolo_gcc_from_stft": (C.c_int, [_P, C.c_int, C.c_int64, C.POINTER(FrontendCfg), _P, _P, _P,
                                       C.POINTER(C.c_int64), _P]),
    "adyolo_scaler_partials": (C.c_int, [_P, C.c_int, C.c_int, C.c_int64, _P, _P, _P, _P, _P]),
    "adyolo_label_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "adyolo_label_cells": (C.c_int, [_P, C.c_int64, C.c_int, C.POINTER(GridCfg), _P, C.c_int64, _P, _P, _P, _P]),
    "adyolo_label_rows": (C.c_int, [_P, C.c_int64, C.POINTER(GridCfg), _P, C.c_int64, _P, _P, _P, C.c_int64, _P]),
    "adyolo_label_cells_rows": (C.c_int, [_P, C.c_int64, C.c_int, C.POINTER(GridCfg), _P, C.c_int64, _P, _P, _P, _P, C.c_int64, _P]),
    "adyolo_launch_count": (C.c_longlong, []),
    "adyolo_loss_bad_rows_offset": (C.c_size_t, []),
    "adyolo_assign": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, C.POINTER(GridCfg), _P, _P, _P, _P]),
    "adyolo_loss_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.POINTER(GridCfg)]),
    "adyolo_loss": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, C.POINTER(GridCfg), _P, _P, _P, _P, _P, _P, _P]),
    "adyolo_loss_devcount": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int, C.c_int, C.POINTER(GridCfg), _P, _P, _P, _P]),
    "adyolo_yolo_post": (C.c_int, [_P, C.c_int64, C.POINTER(GridCfg), C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, _P, _P, _P, _P]),
    "adyolo_loss_backward": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(GridCfg), _P, _P, _P, _P]),
    "adyolo_loss_grad_scale": (C.c_int, [_P, C.c_int64, _P, _P]),
    "adyolo_spec_mask": (C.c_int, [_P, C.c_int, C.c_int, C.c_int64, C.c_int, _P, C.c_int, _P, _P]),
}

_lib = None


def lib():
    """The loaded shared library.  Fails loudly when it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(adyolo_b200 has no CPU/PyTorch fallback).")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            if hasattr(L, name):
                fn = getattr(L, name)
                fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def declared_symbols():
    """Every `adyolo_*` function declared in include/adyolo_b200.h (parsed from the header)."""
    import re
    with open(os.path.join(ROOT, "include", "adyolo_b200.h")) as f:
        txt = f.read()
    return sorted(set(re.findall(r"\b(adyolo_[a-z0-9_]+)\s*\(", txt)))


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().adyolo_last_error().decode(errors="replace")
        if rc == -3:
            raise NotImplementedError(f"{what}: {msg}")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def require_cuda(t, what: str):
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError(f"{what}: a CUDA device is required (adyolo_b200 has no CPU fallback)")
    if t is not None and not t.is_cuda:
        raise RuntimeError(f"{what}: tensor must live on a CUDA device")


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
