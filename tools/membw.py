"""HBM bandwidth of this box by access mix (what a write-dominated kernel can be held against):
fill = write only, copy = read + write, sum = read only.  Buffers of 2 GiB (>> L2), CUDA events."""
import json
import torch

dev = torch.device("cuda:0")
n = 1 << 29                                  # 2 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device=dev)
b = torch.empty(n, dtype=torch.float32, device=dev)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


t_fill = timed(lambda: a.fill_(1.0))
t_copy = timed(lambda: b.copy_(a))
t_sum = timed(lambda: a.sum())
out = {"fill_GBps": 4 * n / t_fill / 1e9, "copy_GBps": 8 * n / t_copy / 1e9, "sum_GBps": 4 * n / t_sum / 1e9}
print(json.dumps(out))
