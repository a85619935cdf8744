"""Stand-alone timing of the fused ball-query+group kernel on bench-shaped inputs (GPU box)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from pvn3d_b200 import _ext, synth

dev = torch.device("cuda:0")
B = int(os.environ.get("QG_B", "32"))
frames = synth.make_batch("linemod", B, n_points=12288, config_id=2, lm_obj_id=1)
cloud = torch.from_numpy(np.stack([f.cld_rgb_nrm for f in frames])).to(dev)
peak, kind = bench.load_peaks()
r = bench.roofline_query_group(torch, _ext, dev, B, cloud, peak, kind)
print(json.dumps({k: v for k, v in r.items() if k != "per_launch"}))
for p in r["per_launch"]:
    print(p)
