import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pvn3d_b200 import synth, fixtures
from pvn3d_b200.eval_utils import FramePoseSolver
dev = torch.device("cuda:0")
B = 32
frames = synth.make_batch("linemod", B, n_points=12288, config_id=2, lm_obj_id=1)
st = synth.stack(frames)
d = {k: torch.from_numpy(v).to(dev) for k, v in st.items()}
s = FramePoseSolver(B, 12288, 8, 2, fixtures.mesh_kps_table_lm(1), None, False, device=dev)
for i in range(2):
    s.solve(d["pcld"], d["labels"], d["ctr_of"], d["kp_of"])
torch.cuda.synchronize()
s.flags |= 4
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); s.solve(d["pcld"], d["labels"], d["ctr_of"], d["kp_of"]); e1.record()
torch.cuda.synchronize()
print("solve ms", e0.elapsed_time(e1))
