"""Pack the reference's object-frame keypoint / radius DATA files into pvn3d_b200/fixtures/*.npz.

These are dataset constants, not code: the 8 FPS
keypoints (`farthest.txt`), the 8 bbox corners whose mean is the object centre (`corners.txt`,
reference basic_utils.py:573-595), the per-class radius list (`radius.txt`, common.py:80) and the
class names (`classes.txt`).  Read exactly as the reference reads them: np.loadtxt(..., float32)
for keypoints/corners, np.loadtxt (float64) for the radius list.

Run in the build container (needs /root/reference): python tools/pack_fixtures.py
"""
import os
import sys

import numpy as np

REF = os.environ.get("PVN3D_REFERENCE", "/root/reference/pvn3d")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pvn3d_b200", "fixtures")

LM_OBJ = {'ape': 1, 'benchvise': 2, 'cam': 4, 'can': 5, 'cat': 6, 'driller': 8, 'duck': 9, 'eggbox': 10,
          'glue': 11, 'holepuncher': 12, 'iron': 13, 'lamp': 14, 'phone': 15}  # common.py:92-106


def main():
    os.makedirs(OUT, exist_ok=True)
    ycb_cls = [l.strip() for l in open(os.path.join(REF, "datasets/ycb/dataset_config/classes.txt")) if l.strip()]
    kps = np.stack([np.loadtxt(os.path.join(REF, "datasets/ycb/ycb_object_kps", c, "farthest.txt"), dtype=np.float32)
                    for c in ycb_cls])
    cors = np.stack([np.loadtxt(os.path.join(REF, "datasets/ycb/ycb_object_kps", c, "corners.txt"), dtype=np.float32)
                     for c in ycb_cls])
    radius = np.loadtxt(os.path.join(REF, "datasets/ycb/dataset_config/radius.txt"))  # float64, common.py:80
    np.savez(os.path.join(OUT, "ycb.npz"), classes=np.array(ycb_cls), farthest=kps, corners=cors, radius=radius)
    print("ycb:", kps.shape, cors.shape, radius.shape)

    names, ids, lk, lc = [], [], [], []
    for name, oid in LM_OBJ.items():
        d = os.path.join(REF, "datasets/linemod/lm_obj_kps", name)
        if not os.path.isdir(d):
            continue
        names.append(name); ids.append(oid)
        lk.append(np.loadtxt(os.path.join(d, "farthest.txt"), dtype=np.float32))
        lc.append(np.loadtxt(os.path.join(d, "corners.txt"), dtype=np.float32))
    np.savez(os.path.join(OUT, "linemod.npz"), names=np.array(names), obj_ids=np.array(ids, np.int32),
             farthest=np.stack(lk), corners=np.stack(lc))
    print("linemod:", len(names), np.stack(lk).shape)


if __name__ == "__main__":
    sys.exit(main())
