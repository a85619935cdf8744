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
//   phase 1  the cloud's xyz streams through shared memory in 24 KB tiles (1-D bulk copy by the TMA
//            engine, mbarrier completion); a warp ballots 32 points per step against its CW centres,
//            hits are appended in index order (the reference's first-nsample rule by construction);
//            the common no-hit step costs 3 LDS + ~11 instructions per centre for both radii.
//   phase 2  descriptors are read POINT-MAJOR (feat_pm[B,N,ldf]).  A warp owns 32 consecutive slots:
//            it pulls 64 channels of two neighbours per LDG.128 (each half-warp one fully used
//            256-byte run), parks them in a private shared-memory tile, and writes the channel-major
//            output as 16-byte stores -- 8 lanes cover one 128-byte line of a channel row, a warp
//            instruction four rows -- after a 4x4 register transpose.  Every output byte is written
//            once, every gathered sector is fully used, ~0.015 instructions per byte.
// Algorithmic HBM bytes per launch and scale:
//   B * [ 12 N + 12 M + 4 C N  (reads)  +  4 M S + 4 (3+C) M S  (writes) ].
#include "common.cuh"

namespace pvn3d {
namespace {

constexpr int kQgThreads = 256;
constexpr int kQgWarps = 8;
constexpr int kQgTile = 2048;      // xyz points per shared-memory tile (24 KB)
constexpr int kQgMaxSlots = 2048;  // slots (centres x nsample, both scales) per CTA
constexpr int kQgTrStride = 68;    // floats per row of the per-warp [32 slots][64 channels] tile

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
  static constexpr size_t tile_bytes = kQgTile * 3 * sizeof(float);                        // 24576
  static constexpr size_t rows_bytes = kQgMaxSlots * sizeof(int);                          // 8192
  static constexpr size_t tr_bytes = kQgWarps * 32 * kQgTrStride * sizeof(float);          // 69632
  static constexpr size_t total = tile_bytes + rows_bytes + tr_bytes;
};

__device__ __forceinline__ void qg_append(unsigned hits, int &cnt, int &first, int ns, int *row,
                                          int kbase, unsigned lane) {
  if (cnt == 0) first = kbase + __ffs(hits) - 1;
  const int slot = cnt + __popc(hits & lanemask_lt());
  if (((hits >> lane) & 1u) && slot < ns) row[slot] = kbase + static_cast<int>(lane);
  cnt += __popc(hits);
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
                   (plane % 4 == 0) && (slot_base % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(sc.out) & 15u) == 0);
  float *tr = s_tr + warp * (32 * kQgTrStride);
  const int ngroups = (nslots + 31) / 32;
  for (int g = warp; g < ngroups; g += kQgWarps) {  // warp-uniform
    const int g0 = g * 32;
    const int glive = min(32, nslots - g0);
    if (vec) {
      const int half = lane >> 4, chunk = lane & 15;  // gather: two neighbours per instruction
      const int m8 = lane >> 2, k4 = lane & 3;        // output: slots 4*m8.., channel quad k4
      for (int c0 = 0; c0 < c; c0 += 64) {
        const int quads = min(16, (c - c0) / 4);
        __syncwarp();
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
          const int q = 2 * i + half;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (q < glive && chunk < quads)
            v = __ldg(reinterpret_cast<const float4 *>(feat_b + static_cast<size_t>(rows[g0 + q]) * a.ldf +
                                                       c0 + 4 * chunk));
          *reinterpret_cast<float4 *>(tr + q * kQgTrStride + 4 * chunk) = v;
        }
        __syncwarp();
        if (4 * m8 < glive) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int quad = 4 * j + k4;
            if (quad < quads) {
              const float4 v0 = *reinterpret_cast<const float4 *>(tr + (4 * m8 + 0) * kQgTrStride + 4 * quad);
              const float4 v1 = *reinterpret_cast<const float4 *>(tr + (4 * m8 + 1) * kQgTrStride + 4 * quad);
              const float4 v2 = *reinterpret_cast<const float4 *>(tr + (4 * m8 + 2) * kQgTrStride + 4 * quad);
              const float4 v3 = *reinterpret_cast<const float4 *>(tr + (4 * m8 + 3) * kQgTrStride + 4 * quad);
              float *dst = out_f + static_cast<size_t>(c0 + 4 * quad) * plane + g0 + 4 * m8;
              if (4 * m8 + 4 <= glive) {
                stg_stream4(dst + 0 * plane, make_float4(v0.x, v1.x, v2.x, v3.x));
                stg_stream4(dst + 1 * plane, make_float4(v0.y, v1.y, v2.y, v3.y));
                stg_stream4(dst + 2 * plane, make_float4(v0.z, v1.z, v2.z, v3.z));
                stg_stream4(dst + 3 * plane, make_float4(v0.w, v1.w, v2.w, v3.w));
              } else {  // ragged last quad of slots
                const float e0[4] = {v0.x, v0.y, v0.z, v0.w}, e1[4] = {v1.x, v1.y, v1.z, v1.w},
                            e2[4] = {v2.x, v2.y, v2.z, v2.w};
                for (int e = 0; e < 4; ++e) {
                  dst[e * plane] = e0[e];
                  if (4 * m8 + 1 < glive) dst[e * plane + 1] = e1[e];
                  if (4 * m8 + 2 < glive) dst[e * plane + 2] = e2[e];
                }
              }
            }
          }
        }
      }
    } else {
      // generic rows (odd channel counts / unaligned strides): scalar 32x32 transposes
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
}

template <int CW, bool DUAL>
__global__ void __launch_bounds__(kQgThreads, 2) query_group_kernel(QgArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float *s_tile = reinterpret_cast<float *>(smem_raw);
  int *s_rows = reinterpret_cast<int *>(smem_raw + QgSmem::tile_bytes);
  float *s_tr = reinterpret_cast<float *>(smem_raw + QgSmem::tile_bytes + QgSmem::rows_bytes);
  __shared__ uint64_t s_bar;

  constexpr int TJ = kQgWarps * CW;
  const int b = blockIdx.y;
  const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
  const float *cloud = a.xyz + static_cast<size_t>(b) * a.n * 3;
  const int jc0 = blockIdx.x * TJ;
  const int live_centres = min(TJ, a.m - jc0);
  const int nsa = a.s[0].ns, nsb = DUAL ? a.s[1].ns : 0;
  const float r2a = __fmul_rn(a.s[0].radius, a.s[0].radius);  // ball_query_gpu.cu:22
  const float r2b = DUAL ? __fmul_rn(a.s[1].radius, a.s[1].radius) : 0.f;
  int *rows_a = s_rows;                  // [TJ][nsa]
  int *rows_b = s_rows + TJ * nsa;       // [TJ][nsb]

  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
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

  unsigned phase = 0;
  bool warp_open = static_cast<int>(warp) * CW < live_centres;
  for (int base = 0; base < a.n; base += kQgTile) {
    const int count = min(kQgTile, a.n - base);
    stage_xyz_tile(s_tile, cloud, base, count, &s_bar, phase, true);
    if (warp_open) {
      for (int off = 0; off < count; off += 32) {
        const int kk = off + static_cast<int>(lane);
        const bool in = kk < count;
        const float x = in ? s_tile[kk * 3 + 0] : 0.f;
        const float y = in ? s_tile[kk * 3 + 1] : 0.f;
        const float z = in ? s_tile[kk * 3 + 2] : 0.f;
#pragma unroll
        for (int q = 0; q < CW; ++q) {
          const float d2 = ref_sqdist(cx[q] - x, cy[q] - y, cz[q] - z);
          const unsigned ha = __ballot_sync(0xffffffffu, in && d2 < r2a);
          const unsigned hb = DUAL ? __ballot_sync(0xffffffffu, in && d2 < r2b) : 0u;
          if (ha | hb) {  // rare: a step with a neighbour in it
            const int lc = static_cast<int>(warp) * CW + q;
            if (ha && cnta[q] < nsa) qg_append(ha, cnta[q], firsta[q], nsa, rows_a + lc * nsa, base + off, lane);
            if (DUAL && hb && cntb[q] < nsb)
              qg_append(hb, cntb[q], firstb[q], nsb, rows_b + lc * nsb, base + off, lane);
          }
        }
        if ((off & 255) == 224) {  // every 8 steps: stop once every ball of this warp is full
          bool open = false;
#pragma unroll
          for (int q = 0; q < CW; ++q) open |= (cnta[q] < nsa) || (DUAL && cntb[q] < nsb);
          if (!open) {
            warp_open = false;
            break;
          }
        }
      }
    }
    // barrier: tile consumed by every warp before it is overwritten; the OR tells all threads the
    // same thing -- whether any ball of this CTA is still unfilled
    if (!__syncthreads_or(warp_open ? 1 : 0)) break;
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
