// query_group.cu -- fused ball-query + grouping for sm_100a (the HBM-bound kernel of the path).
//
// Replaces the five launches + two copies the reference spends per scale in
// QueryAndGroup.forward (pvn3d/lib/pointnet2_utils/pointnet2_utils.py:311-321):
//     ball_query -> transpose(xyz) -> group_points(xyz) -> subtract centre -> group_points(feats)
//     -> torch.cat
// with ONE kernel that writes idx[B,M,S] and the concatenated tensor out[B,3+C,M,S] directly.
//
// Data movement (per CTA = 256 consecutive (centre,sample) slots of one cloud):
//   phase 1  the cloud's xyz streams through shared memory in 24 KB tiles (1-D bulk copy by the
//            TMA engine, mbarrier completion); one warp per centre ballots 32 points per step, so
//            hits are appended in index order (the reference's first-nsample rule).
//   phase 2  descriptors are read POINT-MAJOR (feat_pm[B,N,ldf]): a warp pulls one neighbour row
//            per load instruction, 32 consecutive channels = one fully used 128-byte line,
//            transposes 32 slots x 32 channels through a private padded shared-memory tile, and
//            stores 128-byte lines of the channel-major output with streaming stores.
//            Every output byte is written exactly once; every gathered sector is fully used.
// Algorithmic HBM bytes per launch:
//   B * [ 12 N + 12 M + 4 C N  (reads)  +  4 M S + 4 (3+C) M S  (writes) ].
#include "common.cuh"

namespace pvn3d {
namespace {

constexpr int kQgThreads = 256;
constexpr int kQgWarps = 8;
constexpr int kQgTile = 2048;   // xyz points per shared-memory tile
constexpr int kQgMaxCW = 4;     // centres per warp in phase 1
constexpr int kQgSmallC = 16;   // below this many channels the per-slot gather is used

struct QgSmem {
  // dynamic shared memory carve-up (bytes)
  static constexpr size_t tile_bytes = kQgTile * 3 * sizeof(float);        // 24576
  static constexpr size_t rows_bytes = 256 * sizeof(int);                  // 1024
  static constexpr size_t tr_bytes = kQgWarps * 32 * 33 * sizeof(float);   // 33792
  static constexpr size_t total = tile_bytes + rows_bytes + tr_bytes;
};

__global__ void __launch_bounds__(kQgThreads)
query_group_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                   const float *__restrict__ feat_pm, int ldf, int n, int m, int c, float radius,
                   int ns, int tj /* centres per CTA */, int *__restrict__ idx,
                   float *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *s_tile = reinterpret_cast<float *>(smem_raw);
  int *s_rows = reinterpret_cast<int *>(smem_raw + QgSmem::tile_bytes);
  float *s_tr = reinterpret_cast<float *>(smem_raw + QgSmem::tile_bytes + QgSmem::rows_bytes);
  __shared__ uint64_t s_bar;

  const int b = blockIdx.y;
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  const float *cloud = xyz + static_cast<size_t>(b) * n * 3;
  const int jc0 = blockIdx.x * tj;                 // first centre of this CTA
  const int live_centres = min(tj, m - jc0);       // >= 1
  const int nslots = live_centres * ns;            // <= 256
  const float r2 = __fmul_rn(radius, radius);

  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    mbar_fence_init();
  }

  // ---------------- phase 1: ball query, warp `w` owns centres w, w+8, w+16, w+24 of the CTA -----
  float cx[kQgMaxCW], cy[kQgMaxCW], cz[kQgMaxCW];
  int cnt[kQgMaxCW], first[kQgMaxCW];
#pragma unroll
  for (int q = 0; q < kQgMaxCW; ++q) {
    const int lc = static_cast<int>(warp) + q * kQgWarps;  // local centre
    const bool live = lc < live_centres;
    const float *p = new_xyz + (static_cast<size_t>(b) * m + jc0 + (live ? lc : 0)) * 3;
    cx[q] = p[0];
    cy[q] = p[1];
    cz[q] = p[2];
    cnt[q] = live ? 0 : ns;
    first[q] = 0;
  }
  __syncthreads();

  unsigned phase = 0;
  bool warp_open = cnt[0] < ns;  // warps beyond live_centres have nothing to search
  for (int base = 0; base < n; base += kQgTile) {
    const int count = min(kQgTile, n - base);
    stage_xyz_tile(s_tile, cloud, base, count, &s_bar, phase, true);
    if (warp_open) {
      for (int off = 0; off < count; off += 32) {
        const int kk = off + static_cast<int>(lane);
        const bool in = kk < count;
        const float x = in ? s_tile[kk * 3 + 0] : 0.f;
        const float y = in ? s_tile[kk * 3 + 1] : 0.f;
        const float z = in ? s_tile[kk * 3 + 2] : 0.f;
        bool any_open = false;
#pragma unroll
        for (int q = 0; q < kQgMaxCW; ++q) {
          if (cnt[q] < ns) {  // warp-uniform
            const float d2 = ref_sqdist(cx[q] - x, cy[q] - y, cz[q] - z);
            const unsigned hits = __ballot_sync(0xffffffffu, in && d2 < r2);
            if (hits) {
              if (cnt[q] == 0) first[q] = base + off + __ffs(hits) - 1;
              const int slot = cnt[q] + __popc(hits & lanemask_lt());
              if (((hits >> lane) & 1u) && slot < ns)
                s_rows[(static_cast<int>(warp) + q * kQgWarps) * ns + slot] = base + kk;
              cnt[q] += __popc(hits);
            }
            any_open |= cnt[q] < ns;
          }
        }
        if (!any_open) {
          warp_open = false;
          break;
        }
      }
    }
    if (!__syncthreads_or(warp_open ? 1 : 0)) break;
  }
  // pad the rows in shared memory: slots >= cnt repeat the first hit (0 for an empty ball)
  __syncwarp();
#pragma unroll
  for (int q = 0; q < kQgMaxCW; ++q) {
    const int lc = static_cast<int>(warp) + q * kQgWarps;
    if (lc < live_centres) {
      const int filled = min(cnt[q], ns);
      for (int s = filled + static_cast<int>(lane); s < ns; s += 32) s_rows[lc * ns + s] = first[q];
    }
  }
  __syncthreads();

  // ---------------- phase 2: one thread per slot ----------------------------------------------
  const int t = threadIdx.x;
  const bool live_slot = t < nslots;
  const int p = live_slot ? s_rows[t] : 0;
  const size_t slot_base = static_cast<size_t>(jc0) * ns;       // first slot of the CTA in [M*S]
  const size_t plane = static_cast<size_t>(m) * ns;             // slots per channel plane
  float *out_b = out + static_cast<size_t>(b) * (3 + c) * plane + slot_base;

  if (live_slot) {
    if (idx) idx[static_cast<size_t>(b) * plane + slot_base + t] = p;
    // grouped_xyz - new_xyz  (pointnet2_utils.py:313-314)
    const int lc = t / ns;
    const float *ctr = new_xyz + (static_cast<size_t>(b) * m + jc0 + lc) * 3;
    const float *pt = cloud + static_cast<size_t>(p) * 3;
    stg_stream(out_b + 0 * plane + t, __ldg(pt + 0) - __ldg(ctr + 0));
    stg_stream(out_b + 1 * plane + t, __ldg(pt + 1) - __ldg(ctr + 1));
    stg_stream(out_b + 2 * plane + t, __ldg(pt + 2) - __ldg(ctr + 2));
  }
  if (c == 0) return;
  const float *feat_b = feat_pm + static_cast<size_t>(b) * n * ldf;
  float *out_f = out_b + 3 * plane;

  if (c <= kQgSmallC) {
    // few channels (level 1: rgb + normal): every slot walks its own short row
    if (live_slot) {
      const float *row = feat_b + static_cast<size_t>(p) * ldf;
      for (int ch = 0; ch < c; ++ch) stg_stream(out_f + ch * plane + t, __ldg(row + ch));
    }
    return;
  }

  // wide rows: warp-private 32-slot x 32-channel transposes
  float *tr = s_tr + warp * (32 * 33);
  const int wslot0 = static_cast<int>(warp) * 32;
  if (wslot0 >= nslots) return;                       // warp-uniform
  const int wlive = min(32, nslots - wslot0);         // live slots of this warp
  for (int c0 = 0; c0 < c; c0 += 32) {
    const int cw = min(32, c - c0);
    const bool ch_ok = static_cast<int>(lane) < cw;
    // gather: row q of the tile = channels c0.. of neighbour q  (coalesced 128-byte reads)
#pragma unroll 8
    for (int q = 0; q < 32; ++q) {
      const int pq = __shfl_sync(0xffffffffu, p, q);
      float v = 0.f;
      if (ch_ok && q < wlive) v = __ldg(feat_b + static_cast<size_t>(pq) * ldf + c0 + lane);
      tr[lane * 33 + q] = v;  // transposed store: bank = (lane + q) mod 32, conflict-free
    }
    __syncwarp();
    if (static_cast<int>(lane) < wlive) {
      float *dst = out_f + static_cast<size_t>(c0) * plane + wslot0 + lane;
#pragma unroll 8
      for (int cc = 0; cc < cw; ++cc) stg_stream(dst + cc * plane, tr[cc * 33 + lane]);
    }
    __syncwarp();
  }
}

}  // namespace
}  // namespace pvn3d

extern "C" int pvn3d_query_and_group(const float *xyz, const float *new_xyz, const float *feat_pm,
                                     int ldf, int b, int n, int m, int c, float radius,
                                     int nsample, int *idx, float *out, pvn3d_stream_t stream) {
  using namespace pvn3d;
  if (!xyz || !new_xyz || !out || b < 0 || n <= 0 || m < 0 || c < 0 || nsample < 0)
    return PVN3D_ERR_INVALID_ARG;
  if (c > 0 && (!feat_pm || ldf < c)) return PVN3D_ERR_INVALID_ARG;
  if (b == 0 || m == 0 || nsample == 0) return PVN3D_OK;
  if (nsample > 256 || b > 65535) return PVN3D_ERR_UNSUPPORTED;
  int tj = 256 / nsample;
  if (tj > kQgMaxCW * kQgWarps) tj = kQgMaxCW * kQgWarps;
  static PerDeviceOnce once;
  if (once.first_time())
    PVN3D_CUDA_TRY(cudaFuncSetAttribute(query_group_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)QgSmem::total),
                   "query_group smem attr");
  dim3 grid(ceil_div(m, tj), b);
  query_group_kernel<<<grid, kQgThreads, QgSmem::total, as_stream(stream)>>>(
      xyz, new_xyz, feat_pm, ldf, n, m, c, radius, nsample, tj, idx, out);
  return check_launch("query_group_kernel");
}
