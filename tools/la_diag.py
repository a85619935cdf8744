"""per-step durations of the device-resident loop with / without look-ahead (diagnostics)"""
import sys, os, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pvn3d_b200 import synth
from pvn3d_b200.pipeline import FramePipeline

dev = torch.device("cuda:0")
B = 32
frames = synth.make_batch("linemod", B, config_id=2, lm_obj_id=1)
host = synth.stack(frames)
rot = [{k: torch.from_numpy(np.roll(v, 8 * r, axis=0).copy()).to(dev) for k, v in host.items()} for r in range(4)]
for la in (True, False):
    pipe = FramePipeline("linemod", B, device=dev, lm_obj_id=1)
    n = 40
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    for rep in range(2):
        torch.cuda.synchronize()
        for i in range(n):
            ev[i].record()
            d = rot[i % 4]
            pipe.run_device(d["cld_rgb_nrm"], d["pcld"], d["labels"], d["ctr_of"], d["kp_of"],
                            next_cloud=rot[(i + 1) % 4]["cld_rgb_nrm"] if la else None)
        ev[n].record()
        torch.cuda.synchronize()
        ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
        print("lookahead" if la else "serial   ", "rep", rep, "total/step %.3f" % (ev[0].elapsed_time(ev[n]) / n),
              "median %.3f min %.3f max %.3f" % (statistics.median(ts), min(ts), max(ts)), [round(t, 1) for t in ts[:16]])

# host loop (e2e): pinned host batches, uploads inside the loop
pinned = [FramePipeline.pin_batch({k: np.roll(v, 8 * r, axis=0).copy() for k, v in host.items()}) for r in range(4)]
for la in (True, False):
    pipe = FramePipeline("linemod", B, device=dev, lm_obj_id=1)
    n = 40
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    for rep in range(3):
        torch.cuda.synchronize()
        for i in range(n):
            ev[i].record()
            pipe.run_host(pinned[i % 4], pinned[(i + 1) % 4] if la else None)
        ev[n].record()
        torch.cuda.synchronize()
        ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
        print("host lookahead" if la else "host serial   ", "rep", rep, "total/step %.3f" % (ev[0].elapsed_time(ev[n]) / n),
              "median %.3f min %.3f max %.3f" % (statistics.median(ts), min(ts), max(ts)), [round(t, 1) for t in ts[:24]])
