"""The UNMODIFIED reference Python running on this package's drop-in boundary, on the GPU.

oracle/build_ref_ext.sh stages the reference's lib/ + common.py + datasets/ fixtures under the
git-ignored oracle/_ref/py (it travels to the GPU box with the snapshot).  Here

  * the reference `Pointnet2MSG` (pvn3d/lib/pvn3d.py:46-154) -- its own pointnet2_utils.py /
    pointnet2_modules.py / pytorch_utils.py, nothing of this package's mirror -- runs with
    `lib.pointnet2_utils._ext` bound to pvn3d_b200._ext (pointnet2_utils.py:19) and must produce
    the SAME BITS as the same module object on the reference's own compiled `_ext`
    (oracle/_ref/_ext.so) with TF32 off: every index op is bit-exact, so every cuDNN call sees
    identical operands;
  * the reference `cal_frame_poses` / `cal_frame_poses_lm` (pvn3d_eval_utils.py:37-110,156-201),
    executed as written on CUDA tensors with the reference `MeanShiftTorch`, is compared with the
    same call after compat.patch_post_modules() (what demo.py:22,98-119 would run): class ids equal,
    poses within 1e-4.
"""
import numpy as np
import pytest
import torch

from pvn3d_b200 import _ext as our_ext
from pvn3d_b200 import compat, synth, testing

from helpers import load_ref_ext, load_reference_python

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    r = load_reference_python()
    if r is None:
        pytest.skip("oracle/_ref/py not staged (run oracle/build_ref_ext.sh where /root/reference exists)")
    return r


def _pose_close(p, q, tol=1e-4):
    dr = np.linalg.norm(p[:, :3] - q[:, :3])
    dt = np.linalg.norm(p[:, 3] - q[:, 3]) / max(np.linalg.norm(q[:, 3]), 1e-9)
    return dr <= tol * np.sqrt(3) and dt <= tol, (dr, dt)


def test_reference_pointnet2msg_on_dropin_ext_is_bit_identical(cuda_dev, ref):
    ref_ext = load_ref_ext()
    if ref_ext is None:
        pytest.skip("oracle/_ref/_ext.so not built")
    assert ref.pn2_utils._ext is our_ext, "compat.install() must have bound the drop-in at pointnet2_utils.py:19"
    torch.manual_seed(0)
    model = ref.pvn3d.Pointnet2MSG(input_channels=6)
    testing.randomize_bn_(model, 1)
    model = model.to(cuda_dev).eval()
    frames = synth.make_batch("ycb", 2, n_points=12288, config_id=11)
    x = torch.from_numpy(np.stack([f.cld_rgb_nrm for f in frames])).to(cuda_dev)
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            y_ours = model(x)
            ref.pn2_utils._ext = ref_ext                     # the reference's own compiled kernels
            try:
                y_ref = model(x)
            finally:
                ref.pn2_utils._ext = our_ext
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    assert y_ours.shape == (2, 128, 12288)
    assert torch.equal(y_ours, y_ref), float((y_ours - y_ref).abs().max())
    # and the mirror module of this package (same state_dict) gives the same features
    mine = testing.seeded_pointnet2msg(0, 1).to(cuda_dev)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            y_mirror = mine(x)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    assert float((y_mirror - y_ref).abs().max()) <= 1e-3 * float(y_ref.abs().mean())


def test_reference_sa_module_autograd_on_dropin_ext(cuda_dev, ref):
    """train_*.py path: the reference's autograd Functions (pointnet2_utils.py:67-241) call
    gather_points_grad / group_points_grad of the drop-in"""
    ref_ext = load_ref_ext()
    if ref_ext is None:
        pytest.skip("oracle/_ref/_ext.so not built")
    from lib.pointnet2_utils import pointnet2_modules as ref_mod

    torch.manual_seed(1)
    sa = ref_mod.PointnetSAModuleMSG(npoint=64, radii=[0.1, 0.2], nsamples=[8, 16],
                                     mlps=[[6, 16, 32], [6, 16, 32]]).to(cuda_dev).eval()
    xyz = torch.rand(2, 512, 3, device=cuda_dev)
    grads = []
    for ext in (our_ext, ref_ext):
        ref.pn2_utils._ext = ext
        try:
            feat = torch.rand(2, 6, 512, device=cuda_dev, generator=torch.Generator(cuda_dev).manual_seed(3)).requires_grad_(True)
            prev = torch.backends.cudnn.allow_tf32
            torch.backends.cudnn.allow_tf32 = False
            try:
                _, out = sa(xyz, feat)
                out.square().sum().backward()
            finally:
                torch.backends.cudnn.allow_tf32 = prev
            grads.append(feat.grad.clone())
        finally:
            ref.pn2_utils._ext = our_ext
    # scatter-adds accumulate in a different order: float tolerance, not bits
    assert torch.allclose(grads[0], grads[1], rtol=1e-4, atol=1e-5 * float(grads[1].abs().max()))


@pytest.mark.parametrize("shape", ["ycb", "linemod"])
def test_reference_cal_frame_poses_vs_patched(cuda_dev, ref, shape):
    """reference post-processing on CUDA tensors (what demo.py runs) vs the same entry points after
    compat.patch_post_modules()"""
    import importlib

    f = synth.make_frame(shape, n_points=4096, seed=77, lm_obj_id=1 if shape == "linemod" else None)
    pcld = torch.from_numpy(f.pcld).to(cuda_dev)
    mask = torch.from_numpy(f.labels).to(cuda_dev)
    ctr_of = torch.from_numpy(f.ctr_of).to(cuda_dev)
    kp_of = torch.from_numpy(f.kp_of).to(cuda_dev)
    ev = ref.eval_utils
    orig = (ev.cal_frame_poses, ev.cal_frame_poses_lm, ev.MeanShiftTorch, ref.meanshift.MeanShiftTorch)
    if shape == "ycb":
        ids_ref, poses_ref = ev.cal_frame_poses(pcld, mask, ctr_of, kp_of, True, 22, True)
    else:
        poses_ref = ev.cal_frame_poses_lm(pcld, mask, ctr_of, kp_of, True, 2, False, 1)
    compat.patch_post_modules()
    try:
        assert ev.cal_frame_poses is not orig[0]
        if shape == "ycb":
            ids, poses = ev.cal_frame_poses(pcld, mask, ctr_of, kp_of, True, 22, True)
            assert np.array_equal(ids, ids_ref)
        else:
            poses = ev.cal_frame_poses_lm(pcld, mask, ctr_of, kp_of, True, 2, False, 1)
    finally:
        ev.cal_frame_poses, ev.cal_frame_poses_lm, ev.MeanShiftTorch, ref.meanshift.MeanShiftTorch = orig
    assert len(poses) == len(poses_ref)
    for p, q in zip(poses, poses_ref):
        # the reference ran torch CUDA kernels (soft cross-check: their norm / sum may round differently from
        # the CPU kernels the contract is pinned to), still well inside the bar
        ok, err = _pose_close(np.asarray(p, np.float64), np.asarray(q, np.float64))
        assert ok, err
