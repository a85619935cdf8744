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
//   phase 1  the cloud's xyz streams through shared memory in double-buffered 12 KB tiles (1-D bulk
//            copies by the TMA engine, mbarrier completion; the copy of tile i+1 overlaps the scan of
//            tile i); a warp ballots 32 points per step against its CW centres, hits are appended in
//            index order (the reference's first-nsample rule by construction).  Every 128-point block
//            of a tile carries a bounding box; blocks out of reach of all of the warp's open balls
//            are skipped -- exact pruning that costs nothing on shuffled clouds and removes most of
//            the scan on raster-ordered ones (the reference's samplers keep raster order).
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
constexpr int kQgTile = 1024;      // xyz points per shared-memory tile (12 KB), double buffered
constexpr int kQgBlock = 128;      // points per bounding-box block of a tile (8 blocks per tile)
constexpr int kQgMaxSlots = 2048;  // slots (centres x nsample, both scales) per CTA
constexpr int kQgTrStride = 36;    // floats per row of a per-warp [32 slots][32 channels] tile
constexpr int kQgTrFloats = 32 * kQgTrStride;  // one tile; every warp owns two (double buffering)
static_assert(kQgTile / kQgBlock == kQgWarps, "one bounding-box block per warp");
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

template <int CW, bool DUAL>
__global__ void __launch_bounds__(kQgThreads, 2) query_group_kernel(QgArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *s_tile = reinterpret_cast<float *>(smem_raw);  // two tiles of kQgTile points
  int *s_rows = reinterpret_cast<int *>(smem_raw + QgSmem::tile_bytes);
  float *s_tr = reinterpret_cast<float *>(smem_raw + QgSmem::tile_bytes + QgSmem::rows_bytes);
  __shared__ uint64_t s_bar[2];
  __shared__ float s_bbox[kQgTile / kQgBlock][6];  // per 128-point block of the current tile

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
  const float r2skip = fmaxf(r2a, r2b) * 1.0001f + 1e-12f;
  int *rows_a = s_rows;             // [TJ][nsa]
  int *rows_b = s_rows + TJ * nsa;  // [TJ][nsb]

  if (threadIdx.x == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    mbar_fence_init();
  }
  // ---------------- phase 1: warp w scans for centres w*CW .. w*CW+CW-1 of the CTA ----------------
  float cx[CW], cy[CW], cz[CW];
  int cnta[CW], firsta[CW], cntb[CW], firstb[CW];
#pragma unroll
  for (int q = 0; q < CW; ++q) {
    const int lc = static_cast<int>(warp) * CW + q;
    const bool live = lc < live_centres;
    const float *p = a.new_xyz + (static_cast<size_t>(b) * a.m + jc0 + (live ? lc : 0)) * 3;
    cx[q] = p[0];
    cy[q] = p[1];
    cz[q] = p[2];
    cnta[q] = live ? 0 : nsa;  // dead centres count as full
    cntb[q] = live ? 0 : nsb;
    firsta[q] = firstb[q] = 0;
  }
  __syncthreads();

  // tile i lives in buffer i&1; the copy of tile i+1 is in flight while tile i is scanned
  const int ntiles = (a.n + kQgTile - 1) / kQgTile;
  const bool bulk = ((reinterpret_cast<uintptr_t>(cloud) & 15u) == 0);  // tile offsets are 16-B multiples
  auto tile_count = [&](int i) { return min(kQgTile, a.n - i * kQgTile); };
  auto issue_tile = [&](int i) {  // called by thread 0 (bulk) or by everyone (fallback)
    const int cnt = tile_count(i);
    const unsigned bytes = static_cast<unsigned>(cnt) * 12u;
    float *dst = s_tile + (i & 1) * (kQgTile * 3);
    const float *src = cloud + static_cast<size_t>(i) * kQgTile * 3;
    if (bulk && (bytes & 15u) == 0u) {
      if (threadIdx.x == 0) {
        mbar_expect_tx(&s_bar[i & 1], bytes);
        bulk_g2s(dst, src, bytes, &s_bar[i & 1]);
      }
    } else {
      for (int e = threadIdx.x; e < cnt * 3; e += kQgThreads) dst[e] = __ldg(src + e);
    }
  };
  auto wait_tile = [&](int i, unsigned (&ph)[2]) {
    const unsigned bytes = static_cast<unsigned>(tile_count(i)) * 12u;
    if (bulk && (bytes & 15u) == 0u) {
      mbar_wait(&s_bar[i & 1], ph[i & 1]);
      ph[i & 1] ^= 1u;
    }
  };
  unsigned ph[2] = {0u, 0u};
  bool warp_open = static_cast<int>(warp) * CW < live_centres;
  issue_tile(0);
  for (int ti = 0; ti < ntiles; ++ti) {
    const int base = ti * kQgTile;
    const int count = tile_count(ti);
    const float *tile = s_tile + (ti & 1) * (kQgTile * 3);
    if (ti + 1 < ntiles) issue_tile(ti + 1);  // buffer (ti+1)&1 was released by the barrier ending tile ti-1
    wait_tile(ti, ph);
    __syncthreads();  // fallback path: plain stores of tile ti visible (also orders s_bbox reuse)
    // bounding box of every 128-point block: warp w takes block w
    {
      const int blk0 = static_cast<int>(warp) * kQgBlock;
      float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
      for (int k = blk0 + static_cast<int>(lane); k < min(blk0 + kQgBlock, count); k += 32) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const float v = tile[k * 3 + d];
          lo[d] = fminf(lo[d], v);
          hi[d] = fmaxf(hi[d], v);
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
      if (lane < 3) {
        s_bbox[warp][lane] = lo[lane];
        s_bbox[warp][3 + lane] = hi[lane];
      }
    }
    __syncthreads();
    if (warp_open) {
      for (int blk = 0; blk * kQgBlock < count; ++blk) {
        // distance from each centre to the block's box (0 inside); NaN coordinates fail the skip test
        bool reach = false;
#pragma unroll
        for (int q = 0; q < CW; ++q) {
          const bool open_q = (cnta[q] < nsa) || (DUAL && cntb[q] < nsb);
          const float ex = fmaxf(fmaxf(s_bbox[blk][0] - cx[q], cx[q] - s_bbox[blk][3]), 0.f);
          const float ey = fmaxf(fmaxf(s_bbox[blk][1] - cy[q], cy[q] - s_bbox[blk][4]), 0.f);
          const float ez = fmaxf(fmaxf(s_bbox[blk][2] - cz[q], cz[q] - s_bbox[blk][5]), 0.f);
          reach |= open_q && !(ex * ex + ey * ey + ez * ez > r2skip);
        }
        if (!reach) continue;  // warp-uniform
        const int off_end = min((blk + 1) * kQgBlock, count);
        for (int off = blk * kQgBlock; off < off_end; off += 32) {
          const int kk = off + static_cast<int>(lane);
          const bool in = kk < count;
          const float x = in ? tile[kk * 3 + 0] : 0.f;
          const float y = in ? tile[kk * 3 + 1] : 0.f;
          const float z = in ? tile[kk * 3 + 2] : 0.f;
#pragma unroll
          for (int q = 0; q < CW; ++q) {
            const float d2 = ref_sqdist(cx[q] - x, cy[q] - y, cz[q] - z);
            const unsigned ha = __ballot_sync(0xffffffffu, in && d2 < r2a);
            const unsigned hb = DUAL ? __ballot_sync(0xffffffffu, in && d2 < r2b) : 0u;
            if (ha | hb) {  // rare: a step with a neighbour in it
              const int lc = static_cast<int>(warp) * CW + q;
              if (ha && cnta[q] < nsa)
                qg_append(ha, cnta[q], firsta[q], nsa, rows_a + lc * nsa, base + off, lane);
              if (DUAL && hb && cntb[q] < nsb)
                qg_append(hb, cntb[q], firstb[q], nsb, rows_b + lc * nsb, base + off, lane);
            }
          }
        }
      }
      bool open = false;
#pragma unroll
      for (int q = 0; q < CW; ++q) open |= (cnta[q] < nsa) || (DUAL && cntb[q] < nsb);
      warp_open = open;
    }
    // barrier: tile consumed by every warp before its buffer is refilled; the OR tells all threads the
    // same thing -- whether any ball of this CTA is still unfilled
    if (!__syncthreads_or(warp_open ? 1 : 0)) {
      // a copy of tile ti+1 may still be in flight: drain it before the buffer is reused / the CTA exits
      if (ti + 1 < ntiles) wait_tile(ti + 1, ph);
      break;
    }
  }
  // pad the rows: slots >= cnt repeat the first hit (0 for an empty ball: torch::zeros, ball_query.cpp:19)
  __syncwarp();
#pragma unroll
  for (int q = 0; q < CW; ++q) {
    const int lc = static_cast<int>(warp) * CW + q;
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

int qg_dispatch(QgArgs &a, int b, bool dual, cudaStream_t st) {
  const int ns_tot = a.s[0].ns + (dual ? a.s[1].ns : 0);
  // centres per warp: as many as fit the row buffer; big clouds amortise the scan over 4 centres
  int cw = 4;
  if (a.n <= 4096) cw = 1;
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
