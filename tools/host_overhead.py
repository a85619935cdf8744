"""host time to ENQUEUE one step of the frame pipeline vs the GPU time of the step (is the loop launch-bound?)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pvn3d_b200 import synth
from pvn3d_b200.pipeline import FramePipeline

dev = torch.device("cuda:0")
for shape, B in (("linemod", 32), ("ycb", 16)):
    frames = synth.make_batch(shape, B, config_id=2 if shape == "linemod" else 3, **({"lm_obj_id": 1} if shape == "linemod" else {}))
    host = synth.stack(frames)
    rot = [{k: torch.from_numpy(np.roll(v, (B // 4) * r, axis=0).copy()).to(dev) for k, v in host.items()} for r in range(4)]
    pipe = FramePipeline(shape, B, device=dev, lm_obj_id=1)
    def step(i):
        d = rot[i % 4]
        pipe.run_device(d["cld_rgb_nrm"], d["pcld"], d["labels"], d["ctr_of"], d["kp_of"], next_cloud=rot[(i + 1) % 4]["cld_rgb_nrm"])
    for i in range(8):
        step(i)
    torch.cuda.synchronize()
    n = 40
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(n):
        step(i)
    e1.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    print(f"{shape} b{B}: host enqueue {t_host / n * 1e3:.3f} ms/step, GPU {e0.elapsed_time(e1) / n:.3f} ms/step")
    # host-only cost when the GPU is the bottleneck is hidden; measure it with the queue drained each step
    ts = []
    for i in range(10):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); step(i); ts.append(time.perf_counter() - t0)
    print(f"   enqueue with an empty queue: median {sorted(ts)[5] * 1e3:.3f} ms/step")
