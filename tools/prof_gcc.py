"""GCC-PHAT tensor-core kernel: check against the numpy oracle, time the MIC pipeline and the kernel alone
(GT_B = clips, default 128); `python tools/prof_gcc.py tc` is the ncu target used for profiles/r01_ncu_gcc_tc.txt."""
import os, sys, subprocess, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def run(mode):
    if mode == "simt":
        os.environ["ADYOLO_GCC_SIMT"] = "1"
    import adyolo_b200 as A
    from adyolo_b200.features import features_mic_batched
    rng = np.random.default_rng(2)
    B, N = 3, 24000 * 2 + 600 * 7          # 87 frames/clip -> 261 frames: tail CTA partly empty
    x = rng.standard_normal((B, N, 4)) * 2000
    x[:, :, 1] = np.roll(x[:, :, 0], 7, axis=1) + rng.standard_normal((B, N)) * 50
    x[1, 5000:9000] = 0
    clips = np.clip(x, -32768, 32767).astype(np.int16)
    out = features_mic_batched(torch.from_numpy(clips).cuda())
    torch.cuda.synchronize()
    np.save(f"gpurun_out/gcc_{mode}.npy", out.cpu().numpy())
    if mode.startswith("tc"):
        from oracle import features_np as F
        o = out.cpu().numpy()
        for b in range(B):
            ref = F.features_mic_stack(clips[b])
            print("clip", b, "gcc max abs err vs oracle", np.abs(o[b, 4:] - ref[4:]).max(), "mel", np.abs(o[b, :4] - ref[:4]).max(), flush=True)
    # timing at config-3 size
    Bc, Nc = int(os.environ.get("GT_B", 128)), 24000 * 5
    audio = torch.randint(-3000, 3000, (Bc, Nc, 4), dtype=torch.int16, device="cuda")
    for _ in range(3): o2 = features_mic_batched(audio)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): o2 = features_mic_batched(audio)
    e1.record(); torch.cuda.synchronize()
    print(mode, "MIC features 128x5s: %.3f ms" % (e0.elapsed_time(e1) / 10), "finite", bool(torch.isfinite(o2).all()), flush=True)
    # the GCC kernel alone on a random spectrum of the same size (985 MB in, 79 MB out)
    import ctypes as C
    from adyolo_b200 import _lib
    from adyolo_b200.features import _cfg, ptr, stream_ptr
    L = _lib.lib(); cfg = _cfg()
    T = Nc // 600
    spec = torch.view_as_complex(torch.randn(Bc, T, 601, 4, 2, device="cuda"))
    g = torch.empty(Bc, 6, T, 64, device="cuda")
    st = (C.c_int64 * 4)(6 * T * 64, T * 64, 64, 1)
    def go():
        rc = L.adyolo_gcc_from_stft(ptr(spec), Bc, T, C.byref(cfg), None, None, ptr(g), st, stream_ptr()); assert rc == 0
    for _ in range(3): go()
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): go()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    # generic-stride path: channels-last output (B, T, 64, 6) must hold the same numbers
    g2 = torch.empty(Bc, T, 64, 6, device="cuda")
    st2 = (C.c_int64 * 4)(T * 64 * 6, 1, 64 * 6, 6)
    rc = L.adyolo_gcc_from_stft(ptr(spec), Bc, T, C.byref(cfg), None, None, ptr(g2), st2, stream_ptr()); assert rc == 0
    torch.cuda.synchronize()
    print("strided == natural:", bool(torch.equal(g2.permute(0, 3, 1, 2), g)), flush=True)
    print(mode, "GCC kernel alone: %.3f ms  (%.0f GB/s algorithmic)" % (ms, (spec.numel() * 8 + g.numel() * 4) / ms / 1e6), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1]); sys.exit(0)
    os.makedirs("gpurun_out", exist_ok=True)
    for m in ("tc",):
        r = subprocess.run(["timeout", "120", sys.executable, __file__, m])
        print(m, "exit", r.returncode, flush=True)


