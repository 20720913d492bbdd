"""How long does the cudaMemsetAsync of the loss workspace take (33 MB at config[1])?"""
import torch
from cuda import cudart
n = 2048000 * 16 + 4096
buf = torch.empty(n, dtype=torch.uint8, device="cuda")
big = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for size in (n, n // 2, 4096):
    ts = []
    for _ in range(20):
        big.fill_(1)                       # evict L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        cudart.cudaMemsetAsync(buf.data_ptr(), 0, size, s)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print("memset %9d B: median %.1f us  min %.1f us" % (size, ts[len(ts) // 2], ts[0]))
