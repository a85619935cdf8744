// metrics.cu -- the callers either side of the hot path, on device:
//
//   seg_argmax   per-point class id = argmax over the segmentation logits -- `_, classes_rgbd =
//                torch.max(pred_rgbd_seg, -1)` (reference pvn3d/demo.py:108, train_*_pvn3d.py eval path) --
//                written as the int32 mask pvn3d_frame_poses_batch consumes, so the pipeline input is the
//                network output and no int64 label tensor is materialised.
//   add / add-s  the two pose-error metrics of the evaluation loop, Basic_Utils.cal_add_cuda /
//                cal_adds_cuda (reference pvn3d/lib/utils/basic_utils.py:617-635).  ADD-S is an N x N
//                nearest-neighbour search like the mean-shift density pass; the reference materialises two
//                [N,N,3] tensors per object for it.
#include "common.cuh"

namespace pvn3d {
namespace {

constexpr int kMetThreads = 256;
constexpr int kMetTile = 2048;  // transformed mesh points per shared-memory tile (24 KB)

// first maximal index, like torch.max(dim=-1) (ascending scan, strict '>'; NaN never wins)
__global__ void seg_argmax_kernel(const float *__restrict__ logits, long long rows, int n_cls,
                                  int *__restrict__ labels) {
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= rows) return;
  const float *row = logits + p * n_cls;
  float best = row[0];
  int bi = 0;
  for (int c = 1; c < n_cls; ++c) {
    const float v = row[c];
    if (v > best || (best != best && v == v)) {  // a leading NaN is replaced by the first number
      best = v;
      bi = c;
    }
  }
  labels[p] = bi;
}

__device__ __forceinline__ float3 rt_apply(const float *rt, float x, float y, float z) {
  // p @ R^T + t  (basic_utils.py:620-621): row-vector times transposed rotation = R p + t
  float3 o;
  o.x = __fmaf_rn(rt[2], z, __fmaf_rn(rt[1], y, rt[0] * x)) + rt[3];
  o.y = __fmaf_rn(rt[6], z, __fmaf_rn(rt[5], y, rt[4] * x)) + rt[7];
  o.z = __fmaf_rn(rt[10], z, __fmaf_rn(rt[9], y, rt[8] * x)) + rt[11];
  return o;
}

// grid (ceil(P/256), F).  Thread i owns mesh point i under the GROUND-TRUTH pose; the mesh under the
// PREDICTED pose streams through shared memory.  partial[f][blk] = (sum_i |pd_i - gt_i|, sum_i min_j |pd_j - gt_i|)
__global__ void __launch_bounds__(kMetThreads) add_adds_kernel(const float *__restrict__ pred_rt,
                                                               const float *__restrict__ gt_rt,
                                                               const float *__restrict__ p3ds, int p,
                                                               float2 *__restrict__ partial) {
  __shared__ float4 s_pd[kMetTile];
  __shared__ float s_rt[24];
  __shared__ float2 s_red[kMetThreads / 32];
  const int f = blockIdx.y, t = threadIdx.x;
  if (t < 12) s_rt[t] = pred_rt[f * 12 + t];
  else if (t < 24) s_rt[t] = gt_rt[f * 12 + t - 12];
  __syncthreads();
  const int i = blockIdx.x * kMetThreads + t;
  const bool live = i < p;
  float3 g = make_float3(0.f, 0.f, 0.f), own = g;
  if (live) {
    const float x = p3ds[i * 3 + 0], y = p3ds[i * 3 + 1], z = p3ds[i * 3 + 2];
    g = rt_apply(s_rt + 12, x, y, z);
    own = rt_apply(s_rt, x, y, z);
  }
  float best = __int_as_float(0x7f800000);
  for (int base = 0; base < p; base += kMetTile) {
    const int n = min(kMetTile, p - base);
    __syncthreads();
    for (int q = t; q < n; q += kMetThreads) {
      const float3 v = rt_apply(s_rt, p3ds[(base + q) * 3 + 0], p3ds[(base + q) * 3 + 1], p3ds[(base + q) * 3 + 2]);
      s_pd[q] = make_float4(v.x, v.y, v.z, 0.f);
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < n; ++j) {
      const float4 v = s_pd[j];
      best = fminf(best, torch_sqnorm(v.x - g.x, v.y - g.y, v.z - g.z));
    }
  }
  float2 acc = make_float2(0.f, 0.f);
  if (live) {
    acc.x = __fsqrt_rn(torch_sqnorm(own.x - g.x, own.y - g.y, own.z - g.z));  // ADD term
    acc.y = __fsqrt_rn(best);                                                  // ADD-S term
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
  }
  if ((t & 31) == 0) s_red[t >> 5] = acc;
  __syncthreads();
  if (t == 0) {
    float2 s = s_red[0];
    for (int w = 1; w < kMetThreads / 32; ++w) { s.x += s_red[w].x; s.y += s_red[w].y; }
    partial[static_cast<size_t>(f) * gridDim.x + blockIdx.x] = s;
  }
}

// fixed-order sum of the per-block partials (bitwise reproducible) and the mean
__global__ void add_adds_finish_kernel(const float2 *__restrict__ partial, int blocks, int p, int nfit,
                                       float *__restrict__ add, float *__restrict__ adds) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nfit) return;
  double sa = 0.0, ss = 0.0;
  for (int b = 0; b < blocks; ++b) {
    sa += partial[static_cast<size_t>(f) * blocks + b].x;
    ss += partial[static_cast<size_t>(f) * blocks + b].y;
  }
  if (add) add[f] = static_cast<float>(sa / p);
  if (adds) adds[f] = static_cast<float>(ss / p);
}

}  // namespace
}  // namespace pvn3d

using namespace pvn3d;

extern "C" int pvn3d_seg_argmax(const float *logits, long long rows, int n_cls, int *labels,
                                pvn3d_stream_t stream) {
  if (!logits || !labels || rows < 0 || n_cls < 1) return PVN3D_ERR_INVALID_ARG;
  if (rows == 0) return PVN3D_OK;
  if (rows > 0x7fffffffll * 256) return PVN3D_ERR_UNSUPPORTED;
  seg_argmax_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, as_stream(stream)>>>(logits, rows, n_cls,
                                                                                             labels);
  return check_launch("seg_argmax_kernel");
}

extern "C" size_t pvn3d_pose_add_adds_workspace_bytes(int n_poses, int n_points) {
  if (n_poses < 0 || n_points < 0) return 0;
  return static_cast<size_t>(n_poses > 0 ? n_poses : 1) * ceil_div(n_points > 0 ? n_points : 1, kMetThreads) *
             sizeof(float2) + 256;
}

extern "C" int pvn3d_pose_add_adds(const float *pred_rt, const float *gt_rt, int n_poses, const float *p3ds,
                                   int n_points, float *add, float *adds, void *workspace,
                                   size_t workspace_bytes, pvn3d_stream_t stream) {
  if (!pred_rt || !gt_rt || !p3ds || (!add && !adds) || !workspace || n_poses < 0 || n_points <= 0)
    return PVN3D_ERR_INVALID_ARG;
  if (n_poses == 0) return PVN3D_OK;
  if (n_poses > 65535) return PVN3D_ERR_UNSUPPORTED;
  if (workspace_bytes < pvn3d_pose_add_adds_workspace_bytes(n_poses, n_points)) return PVN3D_ERR_WORKSPACE;
  if (reinterpret_cast<uintptr_t>(workspace) & 7u) return PVN3D_ERR_INVALID_ARG;
  const int blocks = ceil_div(n_points, kMetThreads);
  float2 *partial = static_cast<float2 *>(workspace);
  add_adds_kernel<<<dim3(blocks, n_poses), kMetThreads, 0, as_stream(stream)>>>(pred_rt, gt_rt, p3ds, n_points,
                                                                                partial);
  int rc = check_launch("add_adds_kernel");
  if (rc != PVN3D_OK) return rc;
  add_adds_finish_kernel<<<ceil_div(n_poses, 64), 64, 0, as_stream(stream)>>>(partial, blocks, n_points, n_poses, add,
                                                                              adds);
  return check_launch("add_adds_finish_kernel");
}
