"""kernel timeline (CUPTI through torch.profiler) of a few steady-state steps of the frame pipeline with look-ahead:
per-stream busy time, when the sampling of batch i+1 starts / ends inside step i, gaps on the main stream"""
import sys, os, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from pvn3d_b200 import synth
from pvn3d_b200.pipeline import FramePipeline

dev = torch.device("cuda:0")
shape, B = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("linemod", 32)
frames = synth.make_batch(shape, B, config_id=2 if shape == "linemod" else 3, **({"lm_obj_id": 1} if shape == "linemod" else {}))
host = synth.stack(frames)
rot = [{k: torch.from_numpy(np.roll(v, (B // 4) * r, axis=0).copy()).to(dev) for k, v in host.items()} for r in range(4)]
pipe = FramePipeline(shape, B, device=dev, lm_obj_id=1)
marks = []
def step(i):
    d = rot[i % 4]
    pipe.run_device(d["cld_rgb_nrm"], d["pcld"], d["labels"], d["ctr_of"], d["kp_of"], next_cloud=rot[(i + 1) % 4]["cld_rgb_nrm"])
for i in range(10):
    step(i)
torch.cuda.synchronize()
N = 6
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(N):
        step(i)
    torch.cuda.synchronize()
ev = []
for e in prof.events():
    if e.device_type is not None and "cuda" in str(e.device_type).lower() and e.time_range is not None:
        dur = e.device_time if hasattr(e, "device_time") else e.cuda_time
        ev.append((e.time_range.start, dur, e.name))
ev.sort()
t0 = ev[0][0]
def short(n):
    for key in ("mlp_layer_kernel", "fps_regs_kernel", "ball_scan", "three_nn", "ms_", "transpose", "gather_xyz", "sa_factor", "sa_centre", "nn_", "Memcpy", "Memset"):
        if key in n:
            return key
    return n.split("(")[0][-28:]
span = ev[-1][0] + ev[-1][1] - t0
print(f"{shape} b{B}: {len(ev)} device activities over {span / 1e3:.3f} ms = {span / N / 1e3:.3f} ms/step")
busy = collections.Counter(); cnt = collections.Counter()
for s, d, n in ev:
    busy[short(n)] += d; cnt[short(n)] += 1
for k, v in busy.most_common(14):
    print(f"   {k:28s} {v / N / 1e3:7.3f} ms/step busy  ({cnt[k] / N:.0f} launches/step)")
# the mlp kernels are serial on the main stream: gaps between consecutive ones, and where the fps kernels sit
mlp = [(s - t0, d) for s, d, n in ev if "mlp_layer_kernel" in n]
fps = [(s - t0, d) for s, d, n in ev if "fps_regs_kernel" in n]
per = len(mlp) // N
for st in range(1, N - 1):
    m = mlp[st * per:(st + 1) * per]
    a, b = m[0][0], m[-1][0] + m[-1][1]
    gaps = [m[i + 1][0] - (m[i][0] + m[i][1]) for i in range(per - 1)]
    f = [x for x in fps if a - 1500 <= x[0] <= b]
    print(f" step {st}: first mlp at {a / 1e3:.3f} ms, mlp span {(b - a) / 1e3:.3f} ms, sum of durations {sum(d for _, d in m) / 1e3:.3f}, gaps {sum(gaps) / 1e3:.3f} "
          f"(max {max(gaps):.0f} us); fps kernels in window: " + ", ".join(f"[{(x[0] - a) / 1e3:+.2f}..{(x[0] + x[1] - a) / 1e3:+.2f}]" for x in f))
    print("    mlp durations us:", [round(d) for _, d in m])
json.dump([(s - t0, d, n[:80]) for s, d, n in ev], open("gpurun_out/timeline.json", "w"))
