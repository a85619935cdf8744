import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pvn3d_b200 import synth
from pvn3d_b200.meanshift import MeanShiftTorch
dev = torch.device("cuda:0")
B = 32
frames = synth.make_batch("linemod", B, n_points=12288, config_id=2, lm_obj_id=1)
clouds = []
for f in frames:
    sel = f.labels == 1
    clouds.append(torch.from_numpy(f.pcld[sel] - f.ctr_of[0][sel]))
    for k in range(8):
        clouds.append(torch.from_numpy(f.pcld[sel] - f.kp_of[k][sel]))
counts = [c.shape[0] for c in clouds]
total = sum(counts)
pts = torch.zeros(total, 4)
pts[:, :3] = torch.cat(clouds)
pts = pts.to(dev)
starts = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int32)
fs = torch.from_numpy(starts).to(dev); fc = torch.tensor(counts, dtype=torch.int32, device=dev)
ms = MeanShiftTorch(0.08)
for i in range(2):
    ms.fit_segments(pts, fs, fc, want_labels=False)
torch.cuda.synchronize()
ms.debug_timing = True
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ctr, _, _, _ = ms.fit_segments(pts, fs, fc, want_labels=False); e1.record()
torch.cuda.synchronize()
print("fits", len(counts), "points", total, "total ms", e0.elapsed_time(e1), "iters min/mean/max", ctr[:,3].min().item(), ctr[:,3].mean().item(), ctr[:,3].max().item())
ws, off = ms._last_ws
cfg = ws[off:off + 8192].view(torch.int32).cpu().numpy()
top, done, seeds, tiles = cfg[16:76], cfg[76:136], cfg[136:196], cfg[196:256]
for p in range(60):
    if top[p] == 0: break
    nxt = top[p + 1] if p + 1 < 60 and top[p + 1] else done[p]
    print(f"phase {p:2d} seeds {seeds[p]:7d} tiles {tiles[p]:5d}  tiles {done[p]-top[p]:6d} us   sync+top {nxt-done[p]:5d} us")
