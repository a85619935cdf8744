"""Hot path A on the library's own kernels end to end: FusedPointnet2MSG.

Takes a Pointnet2MSG module (this package's mirror or the reference's -- same state_dict layout),
folds every Conv2d(1x1, no bias) + BatchNorm2d (eval) pair into a TF32-rounded, zero-padded weight
matrix + bias, and runs `Pointnet2MSG.forward` (reference pvn3d/lib/pvn3d.py:126-154) as:

  per SA level : furthest_point_sampling -> (per scale) ball_query -> first SharedMLP layer with the
                 grouping fused into the tensor-core operand producer (pvn3d_mlp_sa_first) -> middle
                 layer (pvn3d_mlp_dense) -> last layer with ReLU + max-pool over nsample fused into the
                 epilogue, written straight into the level's point-major feature table
  per FP level : three_nn -> inverse-distance weights -> first layer with three_interpolate + concat
                 fused into the producer (pvn3d_mlp_fp_first) -> second layer
  last         : [B,N,128] -> [B,128,N] (the layout the reference returns)

The grouped tensors [B,3+C,M,S] and the interpolated tensors [B,C,n] are never materialised; no
cuDNN / cuBLAS / ATen kernel runs in this path.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Tuple

import torch

from . import _ext, _lib
from ._lib import check, ptr
from .pointnet2 import SA_SPEC


def tf32_round(x: torch.Tensor) -> torch.Tensor:
    """round-to-nearest (ties away) to TF32's 10-bit mantissa, like cvt.rna.tf32.f32"""
    i = x.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def fold_conv_bn(layer: torch.nn.Module) -> Tuple[torch.Tensor, torch.Tensor]:
    """layer = Sequential(conv [, normlayer.bn], activation) of SharedMLP (pytorch_utils.py:80-134).
    Returns (W [N,K], bias [N]) with the eval-mode BatchNorm folded in."""
    w = layer.conv.weight.detach().double().flatten(1)
    n = w.size(0)
    bias = layer.conv.bias.detach().double() if layer.conv.bias is not None else torch.zeros(n, dtype=torch.float64, device=w.device)
    if hasattr(layer, "normlayer"):
        bn = layer.normlayer.bn
        scale = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
        w = w * scale[:, None]
        bias = (bias - bn.running_mean.detach().double()) * scale + bn.bias.detach().double()
    return w.float(), bias.float()


class PackedLayer:
    """one folded layer in the layout pvn3d_mlp_* expects"""

    def __init__(self, w: torch.Tensor, bias: torch.Tensor, k_valid_prev_pad: int | None = None):
        n, k = w.shape
        self.n, self.k = n, k
        self.n_pad = (n + 15) // 16 * 16
        k_in = k if k_valid_prev_pad is None else k_valid_prev_pad     # width of the producing activation
        self.k_pad = (max(k, k_in) + 31) // 32 * 32
        wp = torch.zeros((self.n_pad, self.k_pad), dtype=torch.float32, device=w.device)
        wp[:n, :k] = w
        self.w = tf32_round(wp).contiguous()
        self.bias = torch.zeros((self.n_pad,), dtype=torch.float32, device=w.device)
        self.bias[:n] = bias


MLP_RELU, MLP_ROUND_OUT, MLP_A_TF32, MLP_OUT_CN = 1, 2, 4, 16      # include/pvn3d_b200.h PVN3D_MLP_*


def _flags(relu, round_out=False, a_tf32=False, reserve=0):
    """PVN3D_MLP_* flags; reserve = SMs left to concurrent kernels (PVN3D_MLP_RESERVE_SMS)"""
    return ((MLP_RELU if relu else 0) | (MLP_ROUND_OUT if round_out else 0) | (MLP_A_TF32 if a_tf32 else 0)
            | ((int(reserve) & 0xFF) << 8))


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def mlp_dense(a2d: torch.Tensor, layer: PackedLayer, relu=True, pool=0, out=None, col0=0, round_out=False,
              a_tf32=False, reserve=0):
    """a2d [rows, lda] point-major activations (all lda columns valid or zero).  a_tf32: a2d came out of
    a layer run with round_out=True (values already TF32) -> asynchronous copy path."""
    lib = _lib.load()
    rows, lda = a2d.shape
    if out is None:
        out = torch.empty((rows // pool if pool else rows, layer.n_pad), dtype=torch.float32, device=a2d.device)
    with torch.cuda.device(a2d.device):
        rc = lib.pvn3d_mlp_dense(ptr(a2d), lda, lda, rows, ptr(layer.w), ptr(layer.bias), layer.k_pad, layer.n_pad,
                                 _flags(relu, round_out, a_tf32, reserve), pool, ptr(out), out.size(-1), col0, _stream(a2d.device))
    check(rc, "pvn3d_mlp_dense")
    return out


def mlp_sa_first(xyz, new_xyz, feat_pm, ldf, c_feat, idx, layer: PackedLayer, relu=True, pool=0, out=None, col0=0,
                 round_out=False, reserve=0, feat_tf32=False):
    lib = _lib.load()
    b, n = xyz.shape[0], xyz.shape[1]
    m, ns = idx.shape[1], idx.shape[2]
    rows = b * m * ns
    if out is None:
        out = torch.empty((rows // pool if pool else rows, layer.n_pad), dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        rc = lib.pvn3d_mlp_sa_first(ptr(xyz), ptr(new_xyz), feat_pm, ldf, c_feat, ptr(idx), b, n, m, ns, ptr(layer.w),
                                    ptr(layer.bias), layer.k_pad, layer.n_pad, _flags(relu, round_out, feat_tf32, reserve), pool, ptr(out),
                                    out.size(-1), col0, _stream(xyz.device))
    check(rc, "pvn3d_mlp_sa_first")
    return out


def mlp_fp_first(known_feat_pm, nn_idx, nn_w, skip_ptr, lds, c1, layer: PackedLayer, relu=True, round_out=False, reserve=0):
    lib = _lib.load()
    b, m_known, c2 = known_feat_pm.shape
    n_unknown = nn_idx.shape[1]
    out = torch.empty((b * n_unknown, layer.n_pad), dtype=torch.float32, device=known_feat_pm.device)
    with torch.cuda.device(known_feat_pm.device):
        rc = lib.pvn3d_mlp_fp_first(ptr(known_feat_pm), c2, ptr(nn_idx), ptr(nn_w), skip_ptr, lds, c1, b, n_unknown,
                                    m_known, ptr(layer.w), ptr(layer.bias), layer.k_pad, layer.n_pad,
                                    _flags(relu, round_out, reserve=reserve), ptr(out), out.size(-1), 0,
                                    _stream(known_feat_pm.device))
    check(rc, "pvn3d_mlp_fp_first")
    return out


def sa_factor_table(xyz: torch.Tensor, feat_ptr: int, ldf: int, c_feat: int, k_pad: int) -> torch.Tensor:
    """[B,n,3] coordinates + point-major descriptors -> [B*n, k_pad] rows [tf32(f) | hi(x) | lo(x) | 0] (see
    pvn3d_sa_factor_table in include/pvn3d_b200.h)"""
    lib = _lib.load()
    rows = xyz.size(0) * xyz.size(1)
    out = torch.empty((rows, k_pad), dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        rc = lib.pvn3d_sa_factor_table(ptr(xyz), feat_ptr, ldf, c_feat, rows, k_pad, ptr(out), _stream(xyz.device))
    check(rc, "pvn3d_sa_factor_table")
    return out


def sa_centre_term(new_xyz: torch.Tensor, wx: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """V[i] = Wx . c_i - bias for every sampled centre: new_xyz [B,m,3], wx [n_pad,3], bias [n_pad] -> [B*m, n_pad]"""
    lib = _lib.load()
    rows = new_xyz.size(0) * new_xyz.size(1)
    n_pad = wx.size(0)
    out = torch.empty((rows, n_pad), dtype=torch.float32, device=new_xyz.device)
    with torch.cuda.device(new_xyz.device):
        rc = lib.pvn3d_sa_centre_term(ptr(new_xyz), ptr(wx), ptr(bias), rows, n_pad, ptr(out), _stream(new_xyz.device))
    check(rc, "pvn3d_sa_centre_term")
    return out


def mlp_sa_fact(u: torch.Tensor, v: torch.Tensor, idx: torch.Tensor, n: int, layer: PackedLayer, relu=True, pool=0,
                out=None, col0=0, round_out=False, reserve=0):
    """second layer of a factored SA scale: rows relu(U[idx] - V) -> layer (pvn3d_mlp_sa_fact)"""
    lib = _lib.load()
    b, m, ns = idx.shape
    rows = b * m * ns
    if out is None:
        out = torch.empty((rows // pool if pool else rows, layer.n_pad), dtype=torch.float32, device=u.device)
    with torch.cuda.device(u.device):
        rc = lib.pvn3d_mlp_sa_fact(ptr(u), ptr(v), u.size(-1), u.size(-1), ptr(idx), b, n, m, ns, ptr(layer.w), ptr(layer.bias),
                                   layer.k_pad, layer.n_pad, _flags(relu, round_out, reserve=reserve), pool, ptr(out),
                                   out.size(-1), col0, _stream(u.device))
    check(rc, "pvn3d_mlp_sa_fact")
    return out


def mlp_fp_fact(p: torch.Tensor, s_: torch.Tensor, nn_idx: torch.Tensor, nn_w: torch.Tensor, m_known: int,
                layer: PackedLayer, relu=True, round_out=False, reserve=0, out_cn=False):
    """second layer of a factored FP module: rows relu(sum_t w_t P[idx_t] + S) -> layer (pvn3d_mlp_fp_fact).
    out_cn: return [b, n_pad, n_unknown] (channel-major frames, PVN3D_MLP_OUT_CN) instead of [b * n_unknown, n_pad]"""
    lib = _lib.load()
    b, n_unknown = nn_idx.shape[0], nn_idx.shape[1]
    ld = p.size(-1)
    shape = (b, layer.n_pad, n_unknown) if out_cn else (b * n_unknown, layer.n_pad)
    out = torch.empty(shape, dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        rc = lib.pvn3d_mlp_fp_fact(ptr(p), ptr(s_), ld, ld, ptr(nn_idx), ptr(nn_w), b, n_unknown, m_known, ptr(layer.w),
                                   ptr(layer.bias), layer.k_pad, layer.n_pad,
                                   _flags(relu, round_out, reserve=reserve) | (MLP_OUT_CN if out_cn else 0),
                                   ptr(out), layer.n_pad, 0, _stream(p.device))
    check(rc, "pvn3d_mlp_fp_fact")
    return out


class LayerChain:
    """the layers of one SharedMLP as the pvn3d_mlp_layer_t array pvn3d_mlp_{sa,fp}_chain take, plus the
    scratch the chained kernel needs (inter-layer tiles of the CTAs, L2-resident)"""

    def __init__(self, layers: List[PackedLayer]):
        self.layers = layers
        self.arr = (_lib.MlpLayer * len(layers))()
        for i, pl in enumerate(layers):
            self.arr[i].w, self.arr[i].bias = pl.w.data_ptr(), pl.bias.data_ptr()
            self.arr[i].k_pad, self.arr[i].n_pad = pl.k_pad, pl.n_pad
        self.n = len(layers)
        self.ws_bytes = int(_lib.load().pvn3d_mlp_chain_workspace_bytes(ctypes.addressof(self.arr), self.n))
        self.ws = torch.empty((self.ws_bytes + 16,), dtype=torch.uint8, device=layers[0].w.device)
        self.ws_ptr = (self.ws.data_ptr() + 15) // 16 * 16

    @property
    def ptr(self) -> int:
        return ctypes.addressof(self.arr)


def mlp_sa_chain(xyz, new_xyz, feat_pm, ldf, c_feat, idx, chain: LayerChain, pool=0, out=None, col0=0, reserve=0):
    """a whole SA-scale SharedMLP (QueryAndGroup producer -> layers -> max-pool) in one launch"""
    lib = _lib.load()
    b, n = xyz.shape[0], xyz.shape[1]
    m, ns = idx.shape[1], idx.shape[2]
    rows = b * m * ns
    if out is None:
        out = torch.empty((rows // pool if pool else rows, chain.layers[-1].n_pad), dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        rc = lib.pvn3d_mlp_sa_chain(ptr(xyz), ptr(new_xyz), feat_pm, ldf, c_feat, ptr(idx), b, n, m, ns, chain.ptr, chain.n,
                                    _flags(True, reserve=reserve), pool, ptr(out), out.size(-1), col0, chain.ws_ptr,
                                    chain.ws_bytes, _stream(xyz.device))
    check(rc, "pvn3d_mlp_sa_chain")
    return out


def mlp_fp_chain(known_feat_pm, nn_idx, nn_w, skip_ptr, lds, c1, chain: LayerChain, out=None, reserve=0):
    """a whole FP-module SharedMLP (three_interpolate + concat producer -> layers) in one launch"""
    lib = _lib.load()
    b, m_known, c2 = known_feat_pm.shape
    n_unknown = nn_idx.shape[1]
    if out is None:
        out = torch.empty((b * n_unknown, chain.layers[-1].n_pad), dtype=torch.float32, device=known_feat_pm.device)
    with torch.cuda.device(known_feat_pm.device):
        rc = lib.pvn3d_mlp_fp_chain(ptr(known_feat_pm), c2, ptr(nn_idx), ptr(nn_w), skip_ptr, lds, c1, b, n_unknown, m_known,
                                    chain.ptr, chain.n, _flags(True, reserve=reserve), ptr(out), out.size(-1), 0, chain.ws_ptr,
                                    chain.ws_bytes, _stream(known_feat_pm.device))
    check(rc, "pvn3d_mlp_fp_chain")
    return out


def three_nn_weights(dist2: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    w = torch.empty_like(dist2)
    rows = dist2.numel() // 3
    with torch.cuda.device(dist2.device):
        rc = lib.pvn3d_three_nn_weights(ptr(dist2), rows, ptr(w), _stream(dist2.device))
    check(rc, "pvn3d_three_nn_weights")
    return w


class GeoPlan:
    """the coordinate-only half of one Pointnet2MSG.forward (FusedPointnet2MSG.geometry)"""

    def __init__(self, cloud: torch.Tensor, l_xyz: List[torch.Tensor]):
        self.key = (cloud.data_ptr(), tuple(cloud.shape))
        self.l_xyz = l_xyz                       # xyz of levels 0..4
        self.ball: List[Tuple[torch.Tensor, torch.Tensor]] = []     # per SA level: idx of both radii
        self.nn: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}  # per FP level: (idx [B,n,3], weights [B,n,3])
        self.done = None                          # event recorded by whoever computed the plan on a side stream

    def tensors(self):
        for t in self.l_xyz:
            yield t
        # (ball / nn tables are computed on the consumer's stream: FusedPointnet2MSG.queries)


class FusedPointnet2MSG:
    """Inference engine for Pointnet2MSG on libpvn3d_b200 only (see module docstring)."""

    def __init__(self, model: torch.nn.Module, device="cuda", chain: bool | None = None):
        self.dev = torch.device(device)
        #: chain=True: one launch per SharedMLP (inter-layer tiles stay in L2, DRAM traffic of the MLPs -95 %);
        #: False (default): one launch per layer.  Same bits either way (tests/test_mlp_gpu.py).  Measured on B200
        #: the chained kernel is 11 % SLOWER (5.25 vs 4.72 ms per 32-frame batch): with 13 warps
        #: per SM the layers are latency-bound, not HBM-bound, and the chain adds a dependency per layer.
        #: PVN3D_MLP_CHAIN=1 selects it.
        if chain is None:
            chain = os.environ.get("PVN3D_MLP_CHAIN", "0") == "1"
        self.chain = bool(chain)
        model = model.to(self.dev).eval()
        self.sa: List[List[List[PackedLayer]]] = []
        self.sa_out: List[int] = []
        for li, sa in enumerate(model.SA_modules):
            scales = []
            for mlp in sa.mlps:
                layers = []
                prev_pad = None
                for k, layer in enumerate(mlp):
                    w, bias = fold_conv_bn(layer)
                    if k == 0:   # reference column order is [xyz(3) | features]; the producer emits [features | xyz]
                        w = torch.cat([w[:, 3:], w[:, :3]], dim=1)
                    pl = PackedLayer(w, bias, prev_pad)
                    prev_pad = pl.n_pad
                    layers.append(pl)
                scales.append(layers)
            self.sa.append(scales)
            self.sa_out.append(sum(s[-1].n for s in scales))
            assert all(s[-1].n % 4 == 0 for s in scales)
        self.fp: List[List[PackedLayer]] = []
        for fp in model.FP_modules:
            layers, prev_pad = [], None
            for layer in fp.mlp:
                w, bias = fold_conv_bn(layer)
                pl = PackedLayer(w, bias, prev_pad)
                prev_pad = pl.n_pad
                layers.append(pl)
            self.fp.append(layers)
        #: store the level tables of SA1-3 TF32-rounded so that the next level gathers them with cp.async
        #: (identical features: every reader of those tables rounds on staging; PVN3D_MLP_ROUND_TABLES=0 disables)
        self.round_tables = (not self.chain) and os.environ.get("PVN3D_MLP_ROUND_TABLES", "1") != "0"
        #: evaluate the first layer of every SA scale once per POINT instead of once per (centre, neighbour) pair
        #: (it is linear before its ReLU).  PVN3D_MLP_FACTOR=0 keeps the gather-first layers.
        self.factor = (not self.chain) and os.environ.get("PVN3D_MLP_FACTOR", "1") != "0"
        self.sa_fact = []
        for li, sa in enumerate(model.SA_modules):
            per_scale = []
            for si, mlp in enumerate(sa.mlps):
                w, bias = fold_conv_bn(mlp[0])                    # reference column order [xyz(3) | features]
                wf, wxyz = w[:, 3:], w[:, :3]
                first = PackedLayer(torch.cat([wf, wxyz, wxyz], dim=1), torch.zeros_like(bias))   # [W_f | W_x | W_x], bias in V
                n_pad = first.n_pad
                wx = torch.zeros((n_pad, 3), dtype=torch.float32, device=self.dev)
                wx[: w.size(0)] = tf32_round(wxyz.to(self.dev))
                b1 = torch.zeros((n_pad,), dtype=torch.float32, device=self.dev)
                b1[: w.size(0)] = bias.to(self.dev)
                per_scale.append((first, wx.contiguous(), b1))
            self.sa_fact.append(per_scale)
        # factored FP first layers: W1 = [W_k (known columns) | W_s (skip columns)]
        self.fp_fact = []
        skip_c = [model.SA_modules[0].mlps[0][0].conv.weight.size(1) - 3] + self.sa_out[:3]     # skip widths of FP1..FP4
        for i, fp in enumerate(model.FP_modules):
            w, bias = fold_conv_bn(fp.mlp[0])
            c1 = skip_c[i]
            c2 = w.size(1) - c1
            lk = PackedLayer(w[:, :c2].contiguous(), torch.zeros_like(bias))
            ws = w[:, c2:]
            if i == 0:     # the skip of FP1 is the raw cloud (ld 9): read it through the level-0 factor table [f | hi x | lo x]
                ws = torch.cat([ws, torch.zeros((w.size(0), 6), dtype=ws.dtype, device=ws.device)], dim=1)
            self.fp_fact.append((lk, PackedLayer(ws.contiguous(), bias), c2))
        self.sa_chain = [[LayerChain(layers) for layers in scales] for scales in self.sa] if self.chain else None
        self.fp_chain = [LayerChain(layers) for layers in self.fp] if self.chain else None
        self._marks = None

    def _m(self, family: str) -> None:
        """profile(): close the interval since the previous mark and charge it to `family`"""
        if self._marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._marks.append((family, e))

    @torch.no_grad()
    def profile(self, pointcloud: torch.Tensor, reps: int = 3) -> Dict[str, float]:
        """One instrumented forward per repetition: CUDA events on the launching stream between the kernel
        families (fps / ball / mlp / three_nn / glue); returns the median ms per family per call.  The
        events serialise nothing that is not already serial (one stream)."""
        import statistics

        self.forward(pointcloud)
        runs = []
        for _ in range(reps):
            torch.cuda.synchronize(self.dev)
            self._marks = []
            self._m("start")
            self.forward(pointcloud)
            marks, self._marks = self._marks, None
            torch.cuda.synchronize(self.dev)
            acc: Dict[str, float] = {}
            for (_, e0), (fam, e1) in zip(marks[:-1], marks[1:]):
                acc[fam] = acc.get(fam, 0.0) + e0.elapsed_time(e1)
            runs.append(acc)
        return {k: statistics.median(r.get(k, 0.0) for r in runs) for k in runs[0]}

    @torch.no_grad()
    def sampling(self, pointcloud: torch.Tensor, fps_chunk: int = 0) -> "GeoPlan":
        """The four furthest-point samplings of Pointnet2MSG.forward and the sampled centres (reference
        pointnet2_modules.py:44-53): 3708 dependent arg-max iterations on ONE CTA per frame -- latency-bound
        and narrow (B of the 148 SMs).  They depend on the coordinates only, so FramePipeline runs them for
        batch i+1 on a side stream under the shared MLPs of batch i.
        fps_chunk > 0: sample that many frames per launch, so that a caller who keeps only `fps_chunk` SMs free
        for this stream never has more sampling CTAs pending than free SMs."""
        assert pointcloud.is_cuda and pointcloud.is_contiguous() and pointcloud.dtype == torch.float32
        b = pointcloud.size(0)
        xyz = pointcloud[..., :3].contiguous()
        self._m("glue")
        plan = GeoPlan(pointcloud, [xyz])
        lib = _lib.load()
        for li, (npoint, radii, nsamples, _) in enumerate(SA_SPEC):
            x = plan.l_xyz[-1]
            if fps_chunk and fps_chunk < b:
                fidx = torch.empty((b, npoint), dtype=torch.int32, device=self.dev)
                n = x.size(1)
                with torch.cuda.device(self.dev):
                    for f0 in range(0, b, fps_chunk):
                        nb = min(fps_chunk, b - f0)
                        check(lib.pvn3d_furthest_point_sampling(x.data_ptr() + f0 * n * 12, nb, n, npoint,
                                                                fidx.data_ptr() + f0 * npoint * 4, _stream(self.dev)),
                              "pvn3d_furthest_point_sampling")
            else:
                fidx = _ext.furthest_point_sampling(x, npoint)
            self._m("fps")
            plan.l_xyz.append(_ext.gather_xyz(x, fidx))
            self._m("glue")
        return plan

    @torch.no_grad()
    def queries(self, plan: "GeoPlan") -> "GeoPlan":
        """The wide coordinate-only kernels: ball queries of both radii of every SA level and the 3-NN indices +
        inverse-distance weights of every FP level (pointnet2_modules.py:57-60,183-186).  They fill the machine,
        so they run on the stream of the MLPs (0.9 ms per 32-frame batch)."""
        for li, (npoint, radii, nsamples, _) in enumerate(SA_SPEC):
            plan.ball.append(_ext.ball_query2(plan.l_xyz[li + 1], plan.l_xyz[li], radii, nsamples))   # both radii, one pass
        self._m("ball")
        for i in range(3, -1, -1):
            d2, nn_idx = _ext.three_nn(plan.l_xyz[i], plan.l_xyz[i + 1])
            plan.nn[i] = (nn_idx, three_nn_weights(d2))
        self._m("three_nn")
        return plan

    def geometry(self, pointcloud: torch.Tensor, fps_chunk: int = 0) -> "GeoPlan":
        """everything of the forward pass that depends on the coordinates only"""
        return self.queries(self.sampling(pointcloud, fps_chunk))

    @torch.no_grad()
    def features(self, pointcloud: torch.Tensor, plan: "GeoPlan", reserve_sms: int = 0, reserve_levels: int = 5) -> torch.Tensor:
        """The shared MLPs of all SA / FP levels on a geometry plan -> [B,128,N].  reserve_sms: SMs the persistent
        MLP kernels leave free for kernels of other streams (the next batch's sampling); reserve_levels: the SA levels
        0 .. reserve_levels-1 do so (5: the FP modules too) -- the sampling of the next batch is over well before the
        MLPs are, and the later layers then take the whole machine."""
        b, n0, width = pointcloud.shape
        c0 = width - 3
        rs_all = int(reserve_sms)
        rs = rs_all if reserve_levels > 0 else 0
        if not plan.ball:
            self.queries(plan)
        # level-0 descriptors are columns 3.. of the input rows themselves (point-major already)
        feats: List[Tuple[int, int, int]] = [(pointcloud.data_ptr() + 12, width, c0)]   # (address, ld, channels)
        keep = [pointcloud]
        l_xyz = plan.l_xyz
        table0 = None
        for li, (npoint, radii, nsamples, _) in enumerate(SA_SPEC):
            rs = rs_all if li < reserve_levels else 0
            x, new_xyz = l_xyz[li], l_xyz[li + 1]
            fptr, ldf, c_feat = feats[-1]
            out_l = torch.empty((b, npoint, self.sa_out[li]), dtype=torch.float32, device=self.dev)
            col = 0
            table = None
            if self.factor and len(self.sa[li][0]) >= 2:
                table = sa_factor_table(x, fptr, ldf, c_feat, self.sa_fact[li][0][0].k_pad)   # shared by both scales
                if li == 0:
                    table0 = table
            for si, (idx, ns, layers) in enumerate(zip(plan.ball[li], nsamples, self.sa[li])):
                if table is not None:
                    first, wx, b1 = self.sa_fact[li][si]
                    u = mlp_dense(table, first, relu=False, a_tf32=True, reserve=rs)        # once per point
                    v = sa_centre_term(new_xyz, wx, b1)                                     # once per centre
                    last2 = len(layers) == 2
                    h = mlp_sa_fact(u, v, idx, x.size(1), layers[1], pool=ns if last2 else 0, round_out=not last2 or
                                    (self.round_tables and li < 3), reserve=rs,
                                    out=out_l.view(b * npoint, -1) if last2 else None, col0=col if last2 else 0)
                    if not last2:
                        for mid in layers[2:-1]:
                            h = mlp_dense(h, mid, round_out=True, a_tf32=True, reserve=rs)
                        mlp_dense(h, layers[-1], pool=ns, out=out_l.view(b * npoint, -1), col0=col, a_tf32=True, reserve=rs,
                                  round_out=self.round_tables and li < 3)
                    col += layers[-1].n
                    continue
                if self.chain:
                    mlp_sa_chain(x, new_xyz, fptr, ldf, c_feat, idx, self.sa_chain[li][si], pool=ns,
                                 out=out_l.view(b * npoint, -1), col0=col, reserve=rs)
                    col += layers[-1].n
                    continue
                # intermediates are stored TF32-rounded (what the next layer's operand is anyway); so are the
                # level tables of SA1-3, whose only readers round them anyway (SA gather producers, FP skip
                # columns): their rows then go global -> shared by cp.async (SA4's table feeds the fp32
                # interpolation of FP4 and stays unrounded)
                h = mlp_sa_first(x, new_xyz, fptr, ldf, c_feat, idx, layers[0], round_out=True, reserve=rs,
                                 feat_tf32=self.round_tables and li >= 1)
                for mid in layers[1:-1]:
                    h = mlp_dense(h, mid, round_out=True, a_tf32=True, reserve=rs)
                mlp_dense(h, layers[-1], pool=ns, out=out_l.view(b * npoint, -1), col0=col, a_tf32=True, reserve=rs,
                          round_out=self.round_tables and li < 3)
                col += layers[-1].n
            self._m("mlp")
            feats.append((out_l.data_ptr(), out_l.size(-1), out_l.size(-1)))
            keep.append(out_l)
        # feature propagation, deepest first (pvn3d.py:149-152)
        l_feat = list(keep)            # l_feat[i]: tensor owning level i's descriptors (point-major)
        rs = rs_all if reserve_levels >= 5 else 0
        for i in range(3, -1, -1):
            unknown, known = l_xyz[i], l_xyz[i + 1]
            nn_idx, nn_w = plan.nn[i]
            known_feat = l_feat[i + 1]
            if known_feat.dim() == 2:
                known_feat = known_feat.view(b, known.size(1), -1)
            sptr, lds, c1 = feats[i]
            layers = self.fp[i]
            # FP factoring pays only where the known descriptors are much wider than the layer and the skip is narrow:
            # FP1 (256 + 6 -> 128 at 12288 points): 382 vs 420 us; FP2-4 measured 10-40 % SLOWER
            if self.factor and len(layers) == 2 and i == 0 and table0 is not None:
                lk, ls, c2 = self.fp_fact[i]
                kf2d = known_feat.reshape(-1, known_feat.size(-1))
                assert kf2d.size(-1) == c2
                pk = mlp_dense(kf2d, lk, relu=False, reserve=rs)                                   # once per known point
                skip2d = table0 if i == 0 else keep[i].view(-1, keep[i].size(-1))
                sk = mlp_dense(skip2d, ls, relu=False, reserve=rs, a_tf32=(i == 0) or self.round_tables)
                # the module's output IS the network's: written channel-major ([B,128,N]) straight from the accumulator
                cn = layers[1].n_pad in (128, 256) and layers[1].n == layers[1].n_pad and unknown.size(1) % 32 == 0
                h = mlp_fp_fact(pk, sk, nn_idx, nn_w, known.size(1), layers[1], round_out=False, reserve=rs, out_cn=cn)
                if cn:
                    self._m("mlp")
                    return h
            elif self.chain:
                h = mlp_fp_chain(known_feat, nn_idx, nn_w, sptr, lds, c1, self.fp_chain[i], reserve=rs)
            else:
                h = mlp_fp_first(known_feat, nn_idx, nn_w, sptr, lds, c1, layers[0], round_out=True, reserve=rs)
                for li2, lyr in enumerate(layers[1:]):
                    last = li2 == len(layers) - 2        # level tables stay full fp32
                    h = mlp_dense(h, lyr, round_out=not last, a_tf32=True, reserve=rs)
            l_feat[i] = h.view(b, unknown.size(1), -1)
            feats[i] = (h.data_ptr(), h.size(-1), h.size(-1))
            self._m("mlp")
        out_pm = l_feat[0]
        n_out = self.fp[0][-1].n
        if out_pm.size(-1) != n_out:
            out_pm = out_pm[..., :n_out].contiguous()
        out = _ext.transpose_nc_to_cn(out_pm)
        self._m("glue")
        return out

    @torch.no_grad()
    def forward(self, pointcloud: torch.Tensor) -> torch.Tensor:
        """pointcloud [B,N,3+C] f32 contiguous on device -> features [B,128,N] (as the reference returns)"""
        return self.features(pointcloud, self.geometry(pointcloud))

    __call__ = forward
