// query_group.cu -- fused ball-query + grouping for sm_100a (the HBM-bound kernel of the path).
//
// Replaces the five launches + two copies the reference spends per scale in
// QueryAndGroup.forward (pvn3d/lib/pointnet2_utils/pointnet2_utils.py:311-321):
//     ball_query -> transpose(xyz) -> group_points(xyz) -> subtract centre -> group_points(feats)
//     -> torch.cat
// with ONE kernel that writes idx[B,M,S] and the concatenated tensor out[B,3+C,M,S] directly --
// optionally for the TWO radii of a multi-scale-grouping level at once (same centres, same cloud:
// every squared distance is computed once and compared against both radii).
//
// Per CTA (8 warps, CW centres per warp):
//   phase 1  every 128-point block of the cloud gets a bounding box (and every 8 blocks a super box) in
//            shared memory; then each warp walks the cloud ON ITS OWN -- no tile staging, no CTA
//            barriers -- balloting 32 points per step against its CW centres (hits are appended in
//            index order: the reference's first-nsample rule by construction) and skipping every
//            block whose box is out of reach of the box of its centres.  Exact pruning: free on
//            shuffled clouds, removes ~90 % of the scan on raster-ordered ones (the reference's
//            samplers keep raster order).  The cloud is read through L1/L2 (384 contiguous bytes per
//            step); the TMA-staged variant of the scan lives on in pn2_ops.cu:ball_query_kernel.
//   phase 2  descriptors are read POINT-MAJOR (feat_pm[B,N,ldf]).  A warp owns 32 consecutive slots:
//            cp.async pulls 32 channels of four neighbours per instruction (8 lanes = one fully used
//            128-byte run) into one of two private shared-memory tiles while the previous tile is
//            written out as 16-byte stores -- 8 lanes cover one 128-byte line of a channel row, a
//            warp instruction four rows -- after a 4x4 register transpose.  Every output byte is
//            written once, every gathered sector is fully used.
// Algorithmic HBM bytes per launch and scale:
//   B * [ 12 N + 12 M + 4 C N  (reads)  +  4 M S + 4 (3+C) M S  (writes) ].
#include "common.cuh"

namespace pvn3d {
namespace {

constexpr int kQgThreads = 256;
constexpr int kQgWarps = 8;
constexpr int kQgTile = 1024;      // (sizes the 24 KB box table: 1024 boxes of 6 floats)
constexpr int kQgMaxSlots = 2048;  // slots (centres x nsample, both scales) per CTA
constexpr int kQgTrStride = 36;    // floats per row of a per-warp [32 slots][32 channels] tile
constexpr int kQgTrFloats = 32 * kQgTrStride;  // one tile; every warp owns two (double buffering)
static_assert(2 * kQgTrFloats >= 32 * 33, "scalar fallback transposes through the same buffer");

struct QgScale {
  float radius;
  int ns;
  int *idx;    // [B,M,ns] or null
  float *out;  // [B,3+C,M,ns] or null (idx-only launch)
};
struct QgArgs {
  const float *xyz, *new_xyz, *feat;
  int ldf, n, m, c;
  int blk;  // points per bounding-box block (multiple of 32; at most 896 blocks + their super boxes)
  const float *boxes;  // [B][nblk + nsup][6] from qg_boxes_kernel
  QgScale s[2];
};

struct QgSmem {
  static constexpr size_t tile_bytes = 2 * kQgTile * 3 * sizeof(float);                    // 24576
  static constexpr size_t rows_bytes = kQgMaxSlots * sizeof(int);                          // 8192
  static constexpr size_t tr_bytes = kQgWarps * 2 * kQgTrFloats * sizeof(float);           // 73728
  static constexpr size_t total = tile_bytes + rows_bytes + tr_bytes;
};

__device__ __forceinline__ void qg_append(unsigned hits, int &cnt, int &first, int ns, int *row,
                                          int kbase, unsigned lane) {
  if (cnt == 0) first = kbase + __ffs(hits) - 1;
  const int slot = cnt + __popc(hits & lanemask_lt());
  if (((hits >> lane) & 1u) && slot < ns) row[slot] = kbase + static_cast<int>(lane);
  cnt += __popc(hits);
}

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// write one scale's slots of this CTA: xyz difference channels + descriptor channels
__device__ __forceinline__ void qg_write_scale(const QgArgs &a, const QgScale &sc, int b, int jc0,
                                               int live_centres, const int *rows, float *s_tr) {
  const int t = threadIdx.x;
  const unsigned lane = lane_id(), warp = t >> 5;
  const int ns = sc.ns, c = a.c;
  const int nslots = live_centres * ns;
  const size_t plane = static_cast<size_t>(a.m) * ns;
  const size_t slot_base = static_cast<size_t>(jc0) * ns;
  const float *cloud = a.xyz + static_cast<size_t>(b) * a.n * 3;
  float *out_b = sc.out + static_cast<size_t>(b) * (3 + c) * plane + slot_base;

  for (int s = t; s < nslots; s += kQgThreads) {
    const int p = rows[s];
    if (sc.idx) sc.idx[static_cast<size_t>(b) * plane + slot_base + s] = p;
    if (sc.out) {
      // grouped_xyz - new_xyz  (pointnet2_utils.py:313-314)
      const float *ctr = a.new_xyz + (static_cast<size_t>(b) * a.m + jc0 + s / ns) * 3;
      const float *pt = cloud + static_cast<size_t>(p) * 3;
      stg_stream(out_b + 0 * plane + s, __ldg(pt + 0) - __ldg(ctr + 0));
      stg_stream(out_b + 1 * plane + s, __ldg(pt + 1) - __ldg(ctr + 1));
      stg_stream(out_b + 2 * plane + s, __ldg(pt + 2) - __ldg(ctr + 2));
    }
  }
  if (!sc.out || c == 0) return;
  const float *feat_b = a.feat + static_cast<size_t>(b) * a.n * a.ldf;
  float *out_f = out_b + 3 * plane;
  const bool vec = (c % 4 == 0) && (a.ldf % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.feat) & 15u) == 0) &&
                   (plane % 4 == 0) && (slot_base % 4 == 0) && (ns % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(sc.out) & 15u) == 0);
  float *tr = s_tr + warp * (2 * kQgTrFloats);
  const int ngroups = (nslots + 31) / 32;
  if (vec) {
    // Work items of this warp: (group g of 32 slots, chunk of 32 channels).  The neighbour rows of
    // item i+1 are fetched with cp.async (16 B per lane, 8 lanes = one 128-byte run of a row, four
    // rows per instruction) into the other buffer while item i is transposed and stored.
    const int nchunks = (c + 31) / 32;
    const int my_groups = (ngroups > static_cast<int>(warp)) ? (ngroups - static_cast<int>(warp) + kQgWarps - 1) / kQgWarps : 0;
    const int items = my_groups * nchunks;
    const int sub = lane >> 3, chunk = lane & 7;  // gather: neighbour sub-row, 16-byte chunk
    const int m8 = lane >> 2, k4 = lane & 3;      // store: slots 4*m8.., channel quad
    auto issue = [&](int item, float *buf) {
      const int g0 = (static_cast<int>(warp) + (item / nchunks) * kQgWarps) * 32;
      const int c0 = (item % nchunks) * 32;
      const int glive = min(32, nslots - g0);
      const int quads = min(8, (c - c0) / 4);
      if (chunk < quads) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int q = 4 * i + sub;
          if (q < glive)
            cp_async16(buf + q * kQgTrStride + 4 * chunk,
                       feat_b + static_cast<size_t>(rows[g0 + q]) * a.ldf + c0 + 4 * chunk);
        }
      }
      cp_async_commit();
    };
    if (items > 0) issue(0, tr);
    for (int item = 0; item < items; ++item) {
      float *buf = tr + (item & 1) * kQgTrFloats;
      if (item + 1 < items) {
        issue(item + 1, tr + ((item + 1) & 1) * kQgTrFloats);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
      const int g0 = (static_cast<int>(warp) + (item / nchunks) * kQgWarps) * 32;
      const int c0 = (item % nchunks) * 32;
      const int glive = min(32, nslots - g0);  // multiple of 4 (ns % 4 == 0)
      const int quads = min(8, (c - c0) / 4);
      if (4 * m8 < glive) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int quad = 4 * j + k4;
          if (quad < quads) {
            const float4 v0 = *reinterpret_cast<const float4 *>(buf + (4 * m8 + 0) * kQgTrStride + 4 * quad);
            const float4 v1 = *reinterpret_cast<const float4 *>(buf + (4 * m8 + 1) * kQgTrStride + 4 * quad);
            const float4 v2 = *reinterpret_cast<const float4 *>(buf + (4 * m8 + 2) * kQgTrStride + 4 * quad);
            const float4 v3 = *reinterpret_cast<const float4 *>(buf + (4 * m8 + 3) * kQgTrStride + 4 * quad);
            float *dst = out_f + static_cast<size_t>(c0 + 4 * quad) * plane + g0 + 4 * m8;
            stg_stream4(dst + 0 * plane, make_float4(v0.x, v1.x, v2.x, v3.x));
            stg_stream4(dst + 1 * plane, make_float4(v0.y, v1.y, v2.y, v3.y));
            stg_stream4(dst + 2 * plane, make_float4(v0.z, v1.z, v2.z, v3.z));
            stg_stream4(dst + 3 * plane, make_float4(v0.w, v1.w, v2.w, v3.w));
          }
        }
      }
      __syncwarp();  // buffer (item & 1) is refilled two items later
    }
    return;
  }
  if (c <= 16) {
    // few channels (level 1: rgb + normal): every slot walks its own short row
    for (int s = t; s < nslots; s += kQgThreads) {
      const float *row = feat_b + static_cast<size_t>(rows[s]) * a.ldf;
      for (int ch = 0; ch < c; ++ch) stg_stream(out_f + ch * plane + s, __ldg(row + ch));
    }
    return;
  }
  // generic rows (odd channel counts / unaligned strides): scalar 32x32 transposes
  for (int g = warp; g < ngroups; g += kQgWarps) {  // warp-uniform
    const int g0 = g * 32;
    const int glive = min(32, nslots - g0);
    const int p = (static_cast<int>(lane) < glive) ? rows[g0 + lane] : 0;
    for (int c0 = 0; c0 < c; c0 += 32) {
      const int cw = min(32, c - c0);
      __syncwarp();
      for (int q = 0; q < 32; ++q) {
        const int pq = __shfl_sync(0xffffffffu, p, q);
        float v = 0.f;
        if (static_cast<int>(lane) < cw && q < glive)
          v = __ldg(feat_b + static_cast<size_t>(pq) * a.ldf + c0 + lane);
        tr[lane * 33 + q] = v;
      }
      __syncwarp();
      if (static_cast<int>(lane) < glive) {
        float *dst = out_f + static_cast<size_t>(c0) * plane + g0 + lane;
        for (int cc = 0; cc < cw; ++cc) stg_stream(dst + cc * plane, tr[cc * 33 + lane]);
      }
    }
  }
}

// Bounding boxes of every block of `blk` consecutive points (one warp per block) -- computed once per
// cloud, not once per CTA.  Super boxes (8 blocks) are appended by qg_supboxes_kernel.
__global__ void qg_boxes_kernel(const float *__restrict__ xyz, int n, int blk, int nblk, int nsup,
                                float *__restrict__ boxes) {
  const int b = blockIdx.y;
  const unsigned lane = lane_id();
  const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= nblk) return;
  const float *cloud = xyz + static_cast<size_t>(b) * n * 3;
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  const int k_end = min((g + 1) * blk, n);
  for (int k = g * blk + static_cast<int>(lane); k < k_end; k += 32) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const float v = __ldg(cloud + static_cast<size_t>(k) * 3 + d);
      lo[d] = fminf(lo[d], v);
      hi[d] = fmaxf(hi[d], v);
      if (!(v == v)) { lo[d] = -3.0e38f; hi[d] = 3.0e38f; }  // NaN coordinate: never skip this block
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
      hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
    }
  }
  if (lane == 0) {
    float *o = boxes + (static_cast<size_t>(b) * (nblk + nsup) + g) * 6;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      o[d] = lo[d];
      o[3 + d] = hi[d];
    }
  }
}
__global__ void qg_supboxes_kernel(int nblk, int nsup, float *__restrict__ boxes) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nsup * 6) return;
  float *base = boxes + static_cast<size_t>(b) * (nblk + nsup) * 6;
  const int sb = e / 6, d = e % 6;
  float v = base[(sb * 8) * 6 + d];
  for (int i = 1; i < 8 && sb * 8 + i < nblk; ++i)
    v = d < 3 ? fminf(v, base[(sb * 8 + i) * 6 + d]) : fmaxf(v, base[(sb * 8 + i) * 6 + d]);
  base[(nblk + sb) * 6 + d] = v;
}

// squared distance between two axis-aligned boxes (0 when they overlap); b[0..2] = lo, b[3..5] = hi
__device__ __forceinline__ float qg_box_dist2(const float (&u)[6], const float *b) {
  const float ex = fmaxf(fmaxf(b[0] - u[3], u[0] - b[3]), 0.f);
  const float ey = fmaxf(fmaxf(b[1] - u[4], u[1] - b[4]), 0.f);
  const float ez = fmaxf(fmaxf(b[2] - u[5], u[2] - b[5]), 0.f);
  return ex * ex + ey * ey + ez * ez;
}

template <int CW, bool DUAL>
__global__ void __launch_bounds__(kQgThreads, 2) query_group_kernel(QgArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *s_box = reinterpret_cast<float *>(smem_raw);  // [nblk][6] block boxes, then [nsup][6] super boxes
  int *s_rows = reinterpret_cast<int *>(smem_raw + QgSmem::tile_bytes);
  float *s_tr = reinterpret_cast<float *>(smem_raw + QgSmem::tile_bytes + QgSmem::rows_bytes);
  __shared__ int s_perm[32];  // scan slot -> local centre, sorted by y

  constexpr int TJ = kQgWarps * CW;
  const int b = blockIdx.y;
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  const float *cloud = a.xyz + static_cast<size_t>(b) * a.n * 3;
  const int jc0 = blockIdx.x * TJ;
  const int live_centres = min(TJ, a.m - jc0);
  const int nsa = a.s[0].ns, nsb = DUAL ? a.s[1].ns : 0;
  const float r2a = __fmul_rn(a.s[0].radius, a.s[0].radius);  // ball_query_gpu.cu:22
  const float r2b = DUAL ? __fmul_rn(a.s[1].radius, a.s[1].radius) : 0.f;
  // a block can be skipped when even its bounding box is out of reach of the larger radius; the bound
  // is inflated so that fp32 rounding of the box distance can never hide a true hit
  const float r2max = fmaxf(r2a, r2b);
  const float r2skip = r2max * 1.0001f + 1e-12f;
  int *rows_a = s_rows;             // [TJ][nsa]
  int *rows_b = s_rows + TJ * nsa;  // [TJ][nsb]

  // ---- bounding boxes of the cloud's blocks (a.blk points each) and of groups of 8 blocks -----------
  const int blk = a.blk, nblk = (a.n + blk - 1) / blk, nsup = (nblk + 7) / 8;
  float *s_sup = s_box + nblk * 6;
  {
    const float *src = a.boxes + static_cast<size_t>(b) * (nblk + nsup) * 6;
    for (int e = threadIdx.x; e < (nblk + nsup) * 6; e += kQgThreads) s_box[e] = __ldg(src + e);
  }
  // The CTA's centres are dealt to the warps in order of their y coordinate, so that the CW centres a
  // warp scans for together are close to each other and share the blocks they can skip (FPS order
  // scatters consecutive centres all over the cloud).  Any assignment gives the same output.
  if (warp == 0) {
    const int lc0 = static_cast<int>(lane);
    float key = (lc0 < live_centres) ? a.new_xyz[(static_cast<size_t>(b) * a.m + jc0 + lc0) * 3 + 1]
                                     : __int_as_float(0x7f800000);
    if (!(key == key)) key = __int_as_float(0x7f800000);  // NaN sorts last
    int val = lc0;
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {  // bitonic sort of (key, val) across the warp
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        const float ok = __shfl_xor_sync(0xffffffffu, key, j);
        const int ov = __shfl_xor_sync(0xffffffffu, val, j);
        const bool up = ((lane & k) == 0);
        const bool lower = ((lane & j) == 0);
        const bool less = (ok < key) || (ok == key && ov < val);
        const bool take = (lower == up) ? less : !less && !(ok == key && ov == val);
        if (take) { key = ok; val = ov; }
      }
    }
    s_perm[lane] = val;
  }
  __syncthreads();
  // ---------------- phase 1: warp w scans for the centres in scan slots w*CW .. w*CW+CW-1 ------------
  // No tile staging, no CTA barriers: every warp walks the cloud on its own (L1/L2-resident, 384
  // contiguous bytes per step) and only where the box tests say a neighbour can be.
  float cx[CW], cy[CW], cz[CW];
  int cnta[CW], firsta[CW], cntb[CW], firstb[CW], lcs[CW];
  float ub[6] = {3.0e38f, 3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};  // box of the warp's centres
  bool warp_open = false;
#pragma unroll
  for (int q = 0; q < CW; ++q) {
    const int lc = s_perm[static_cast<int>(warp) * CW + q];
    lcs[q] = lc;
    const bool live = lc < live_centres;
    const float *p = a.new_xyz + (static_cast<size_t>(b) * a.m + jc0 + (live ? lc : 0)) * 3;
    cx[q] = p[0];
    cy[q] = p[1];
    cz[q] = p[2];
    cnta[q] = live ? 0 : nsa;  // dead centres count as full
    cntb[q] = live ? 0 : nsb;
    firsta[q] = firstb[q] = 0;
    if (live) {
      warp_open = true;
      ub[0] = fminf(ub[0], cx[q]); ub[1] = fminf(ub[1], cy[q]); ub[2] = fminf(ub[2], cz[q]);
      ub[3] = fmaxf(ub[3], cx[q]); ub[4] = fmaxf(ub[4], cy[q]); ub[5] = fmaxf(ub[5], cz[q]);
      if (!(cx[q] == cx[q]) || !(cy[q] == cy[q]) || !(cz[q] == cz[q])) {  // NaN centre: no pruning
        ub[0] = ub[1] = ub[2] = -3.0e38f;
        ub[3] = ub[4] = ub[5] = 3.0e38f;
      }
    }
  }
  for (int sb = 0; sb < nsup && warp_open; ++sb) {
    if (qg_box_dist2(ub, s_sup + sb * 6) > r2skip) continue;
    const int g_end = min(sb * 8 + 8, nblk);
    for (int g = sb * 8; g < g_end && warp_open; ++g) {
      if (qg_box_dist2(ub, s_box + g * 6) > r2skip) continue;
      const int k_end = min((g + 1) * blk, a.n);
      for (int off = g * blk; off < k_end; off += 32) {
        const int kk = off + static_cast<int>(lane);
        const bool in = kk < k_end;
        const float *pt = cloud + static_cast<size_t>(in ? kk : off) * 3;
        const float x = __ldg(pt), y = __ldg(pt + 1), z = __ldg(pt + 2);
#pragma unroll
        for (int q = 0; q < CW; ++q) {
          const float d2 = ref_sqdist(cx[q] - x, cy[q] - y, cz[q] - z);
          // one ballot against the larger radius decides the common no-neighbour step
          if (__ballot_sync(0xffffffffu, in && d2 < r2max)) {
            const unsigned ha = __ballot_sync(0xffffffffu, in && d2 < r2a);
            const unsigned hb = DUAL ? __ballot_sync(0xffffffffu, in && d2 < r2b) : 0u;
            if (ha && cnta[q] < nsa)
              qg_append(ha, cnta[q], firsta[q], nsa, rows_a + lcs[q] * nsa, off, lane);
            if (DUAL && hb && cntb[q] < nsb)
              qg_append(hb, cntb[q], firstb[q], nsb, rows_b + lcs[q] * nsb, off, lane);
          }
        }
      }
      bool open = false;
#pragma unroll
      for (int q = 0; q < CW; ++q) open |= (cnta[q] < nsa) || (DUAL && cntb[q] < nsb);
      warp_open = open;  // every ball of this warp is full: nothing further can be appended
    }
  }
  // pad the rows: slots >= cnt repeat the first hit (0 for an empty ball: torch::zeros, ball_query.cpp:19)
  __syncwarp();
#pragma unroll
  for (int q = 0; q < CW; ++q) {
    const int lc = lcs[q];
    if (lc < live_centres) {
      for (int s = min(cnta[q], nsa) + static_cast<int>(lane); s < nsa; s += 32) rows_a[lc * nsa + s] = firsta[q];
      if (DUAL)
        for (int s = min(cntb[q], nsb) + static_cast<int>(lane); s < nsb; s += 32) rows_b[lc * nsb + s] = firstb[q];
    }
  }
  __syncthreads();

  // ---------------- phase 2 ------------------------------------------------------------------------
  qg_write_scale(a, a.s[0], b, jc0, live_centres, rows_a, s_tr);
  if (DUAL) qg_write_scale(a, a.s[1], b, jc0, live_centres, rows_b, s_tr);
}

template <int CW, bool DUAL>
int qg_launch(const QgArgs &a, int b, cudaStream_t st) {
  auto kern = query_group_kernel<CW, DUAL>;
  static PerDeviceOnce once;
  if (once.first_time())
    PVN3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)QgSmem::total),
                   "query_group smem attr");
  dim3 grid(ceil_div(a.m, kQgWarps * CW), b);
  kern<<<grid, kQgThreads, QgSmem::total, st>>>(a);
  return check_launch("query_group_kernel");
}

int qg_dispatch_launch(QgArgs &a, int b, bool dual, cudaStream_t st);

int qg_dispatch(QgArgs &a, int b, bool dual, cudaStream_t st) {
  a.blk = 128;
  while (ceil_div(a.n, a.blk) > 896) a.blk *= 2;  // box table lives in 24 KB of shared memory
  const int nblk = ceil_div(a.n, a.blk), nsup = ceil_div(nblk, 8);
  int rc0 = keep_async_pool_warm();
  if (rc0 != PVN3D_OK) return rc0;
  float *boxes = nullptr;  // stream-ordered scratch: [B][nblk+nsup][6] floats (2.6 KB per 12288-pt cloud)
  PVN3D_CUDA_TRY(cudaMallocAsync(&boxes, sizeof(float) * 6 * static_cast<size_t>(b) * (nblk + nsup), st),
                 "query_group box scratch");
  qg_boxes_kernel<<<dim3(ceil_div(nblk, 8), b), 256, 0, st>>>(a.xyz, a.n, a.blk, nblk, nsup, boxes);
  int rc = check_launch("qg_boxes_kernel");
  if (rc == PVN3D_OK) {
    qg_supboxes_kernel<<<dim3(ceil_div(nsup * 6, 128), b), 128, 0, st>>>(nblk, nsup, boxes);
    rc = check_launch("qg_supboxes_kernel");
  }
  a.boxes = boxes;
  if (rc == PVN3D_OK) rc = qg_dispatch_launch(a, b, dual, st);
  cudaFreeAsync(boxes, st);
  return rc;
}

int qg_dispatch_launch(QgArgs &a, int b, bool dual, cudaStream_t st) {
  const int ns_tot = a.s[0].ns + (dual ? a.s[1].ns : 0);
  // centres per warp: as many as fit the row buffer; big clouds amortise the scan over 4 centres
  int cw = 4;
  if (a.n < 2048) cw = 2;
  if (a.n < 1024) cw = 1;
  while (cw > 1 && kQgWarps * cw * ns_tot > kQgMaxSlots) cw >>= 1;
  if (kQgWarps * cw * ns_tot > kQgMaxSlots) return PVN3D_ERR_UNSUPPORTED;
  if (dual) {
    if (cw == 4) return qg_launch<4, true>(a, b, st);
    if (cw == 2) return qg_launch<2, true>(a, b, st);
    return qg_launch<1, true>(a, b, st);
  }
  if (cw == 4) return qg_launch<4, false>(a, b, st);
  if (cw == 2) return qg_launch<2, false>(a, b, st);
  return qg_launch<1, false>(a, b, st);
}

}  // namespace
}  // namespace pvn3d

using namespace pvn3d;

extern "C" int pvn3d_query_and_group(const float *xyz, const float *new_xyz, const float *feat_pm,
                                     int ldf, int b, int n, int m, int c, float radius,
                                     int nsample, int *idx, float *out, pvn3d_stream_t stream) {
  if (!xyz || !new_xyz || !out || b < 0 || n <= 0 || m < 0 || c < 0 || nsample < 0)
    return PVN3D_ERR_INVALID_ARG;
  if (c > 0 && (!feat_pm || ldf < c)) return PVN3D_ERR_INVALID_ARG;
  if (b == 0 || m == 0 || nsample == 0) return PVN3D_OK;
  if (b > 65535) return PVN3D_ERR_UNSUPPORTED;
  QgArgs a{};
  a.xyz = xyz; a.new_xyz = new_xyz; a.feat = feat_pm; a.ldf = ldf; a.n = n; a.m = m; a.c = c;
  a.s[0] = QgScale{radius, nsample, idx, out};
  return qg_dispatch(a, b, false, as_stream(stream));
}

extern "C" int pvn3d_query_and_group2(const float *xyz, const float *new_xyz, const float *feat_pm,
                                      int ldf, int b, int n, int m, int c, float radius0,
                                      int nsample0, int *idx0, float *out0, float radius1,
                                      int nsample1, int *idx1, float *out1, pvn3d_stream_t stream) {
  if (!xyz || !new_xyz || b < 0 || n <= 0 || m < 0 || c < 0 || nsample0 <= 0 || nsample1 <= 0)
    return PVN3D_ERR_INVALID_ARG;
  if ((!idx0 && !out0) || (!idx1 && !out1)) return PVN3D_ERR_INVALID_ARG;
  if (c > 0 && (out0 || out1) && (!feat_pm || ldf < c)) return PVN3D_ERR_INVALID_ARG;
  if (b == 0 || m == 0) return PVN3D_OK;
  if (b > 65535) return PVN3D_ERR_UNSUPPORTED;
  QgArgs a{};
  a.xyz = xyz; a.new_xyz = new_xyz; a.feat = feat_pm; a.ldf = ldf; a.n = n; a.m = m; a.c = c;
  a.s[0] = QgScale{radius0, nsample0, idx0, out0};
  a.s[1] = QgScale{radius1, nsample1, idx1, out1};
  return qg_dispatch(a, b, true, as_stream(stream));
}
