"""event-timed furthest_point_sampling on bench-shaped clouds (GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pvn3d_b200 import _ext, synth
dev = torch.device("cuda:0")
frames = synth.make_batch("linemod", 32, n_points=12288, config_id=2, lm_obj_id=1)
xyz = torch.from_numpy(np.stack([f.cld_rgb_nrm[:, :3] for f in frames])).contiguous().to(dev)
for n, m in ((12288, 2048), (2048, 1024)):
    x = xyz[:, :n].contiguous()
    for _ in range(3):
        _ext.furthest_point_sampling(x, m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        _ext.furthest_point_sampling(x, m)
    e1.record(); torch.cuda.synchronize()
    print(f"n={n} m={m}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us  ({e0.elapsed_time(e1) / 5 / m * 1e3:.3f} us/iteration)")
