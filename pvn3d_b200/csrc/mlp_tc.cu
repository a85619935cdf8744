// mlp_tc.cu -- the shared-MLP contractions of set abstraction / feature propagation on the 5th-gen
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM), with the PointNet++ data movement fused
// into the operand producer and the activation / max-pool fused into the epilogue.
//
// Replaces, per SharedMLP layer of the reference (pytorch_utils.py:25-50: Conv2d 1x1 (no bias) ->
// BatchNorm2d -> ReLU, plus F.max_pool2d over nsample at pointnet2_modules.py:64-67): one cuDNN conv,
// one cuDNN BN, one clamp kernel, [one max-pool kernel] and -- for the first layer of a scale -- the
// whole materialised grouped tensor [B,3+C,M,S] (61 MB / frame over the eight scales).
//
//   out[p, :] = act( A[p, :] . W^T + bias )            W: BN folded, K-major, TF32-rounded
//
// A-operand producers (warps 0-3, one thread per row of the 128-row tile):
//   DENSE      rows of a point-major activation matrix [P, lda]
//   SA_GATHER  row (b,j,s): [ feat_pm[b, idx[b,j,s], 0:C] | xyz[b,idx] - new_xyz[b,j] | 0 ... ]
//              (QueryAndGroup.forward, pointnet2_utils.py:311-321, never materialised; the weight
//              matrix has its xyz columns moved behind the feature columns to keep rows 16-B aligned)
//   FP_INTERP  row (b,j): [ sum_t w_t * known_feat_pm[b, idx_t, 0:C2] | skip[b, j, 0:C1] | 0 ... ]
//              (three_interpolate + torch.cat of PointnetFPModule.forward, pointnet2_modules.py:188-199)
// Operands are staged in shared memory in the canonical K-major SWIZZLE_128B layout (32 fp32 = 128 B
// per row per stage), one elected thread of warp 4 issues tcgen05.mma (M=128, N<=256, K=8 per
// instruction) and releases stages with tcgen05.commit; the epilogue warps read the accumulator with
// tcgen05.ld (32x32b), add the folded-BN bias, apply ReLU and either store the point-major row or
// reduce over the nsample rows of each centre with a transposing shuffle butterfly (max-pool) and
// store one 128-byte line per centre.
//
// TF32: operands are rounded to nearest (cvt.rna.tf32.f32) when staged, accumulation is fp32 -- the
// precision class of the reference's default cuDNN convolutions (torch.backends.cudnn.allow_tf32).
#include "common.cuh"

namespace pvn3d {
namespace {

constexpr int kMlpBM = 128;
constexpr int kMlpThreads = 160;  // warps 0-3: producers + epilogue; warp 4: TMEM owner + MMA issuer
constexpr int kMlpMaxStages = 4;

enum : int { PRO_DENSE = 0, PRO_SA_GATHER = 1, PRO_FP_INTERP = 2 };
enum : int { EPI_STORE = 0, EPI_MAXPOOL = 1 };

struct MlpArgs {
  // GEMM
  const float *w;     // [n_pad][k_pad]
  const float *bias;  // [n_pad]
  long long rows;     // P
  int k_pad;          // multiple of 32
  int n_pad;          // multiple of 16
  int bn;             // columns per CTA (multiple of 16, <= 256)
  int stages;
  int tmem_cols;      // power of two >= max(32, bn)
  // DENSE
  const float *a;
  int lda;
  int a_cols;  // valid columns of a (multiple of 4); the rest of k_pad reads as zero
  // SA_GATHER
  const float *xyz, *new_xyz, *feat;
  const int *idx;
  int ldf, c_feat, n, m, ns;
  // FP_INTERP
  const float *known_feat, *nn_w, *skip;
  const int *nn_idx;
  int c2, lds, c1, n_unknown, m_known;
  // epilogue
  float *out;
  int ldo, col0, relu;
  int pool;  // nsample of the max-pool epilogue (8, 16 or 32)
};

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ float4 ldg128(const float *p) {
  return __ldg(reinterpret_cast<const float4 *>(p));
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 columns of fp32: thread i of the warp gets row (lane base + i), columns c..c+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
      "%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[32]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
#pragma unroll
  for (int i = 16; i < 32; ++i) v[i] = 0.f;
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4,
// LBO = 1 (unused for swizzled K-major), SBO = 1024 B (8 rows x 128 B), version 1, layout 2 (SW128).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// cute::UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, A and B K-major, M = 128
__device__ __forceinline__ uint32_t instr_desc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(kMlpBM >> 4) << 24);
}

// byte offset of 16-byte chunk `c` (0..7) of row `r` inside a [rows x 128 B] SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_off(int r, int c) {
  return static_cast<uint32_t>(r) * 128u + (static_cast<uint32_t>(c ^ (r & 7)) << 4);
}

// ---- A-operand row sources -------------------------------------------------------------------------
struct RowSrc {
  // DENSE / generic
  const float *row;  // start of this row (DENSE: a + p*lda; SA: feature row; FP: skip row)
  bool live;
  // SA
  float dx, dy, dz;
  // FP
  const float *k1, *k2, *k3;
  float w1, w2, w3;
};

template <int PRO>
__device__ __forceinline__ void row_setup(const MlpArgs &a, long long p, RowSrc &s) {
  s.live = p < a.rows;
  s.row = nullptr;
  if (!s.live) return;
  if (PRO == PRO_DENSE) {
    s.row = a.a + p * a.lda;
  } else if (PRO == PRO_SA_GATHER) {
    const long long per_b = static_cast<long long>(a.m) * a.ns;
    const int b = static_cast<int>(p / per_b);
    const int r = static_cast<int>(p - b * per_b);
    const int j = r / a.ns;
    const int q = a.idx[p];
    const float *pt = a.xyz + (static_cast<size_t>(b) * a.n + q) * 3;
    const float *ct = a.new_xyz + (static_cast<size_t>(b) * a.m + j) * 3;
    s.dx = __ldg(pt) - __ldg(ct);  // grouped_xyz -= new_xyz (pointnet2_utils.py:314)
    s.dy = __ldg(pt + 1) - __ldg(ct + 1);
    s.dz = __ldg(pt + 2) - __ldg(ct + 2);
    s.row = a.c_feat ? a.feat + (static_cast<size_t>(b) * a.n + q) * a.ldf : nullptr;
  } else {
    const int b = static_cast<int>(p / a.n_unknown);
    const int *ii = a.nn_idx + p * 3;
    const float *ww = a.nn_w + p * 3;
    const float *kb = a.known_feat + static_cast<size_t>(b) * a.m_known * a.c2;
    s.k1 = kb + static_cast<size_t>(ii[0]) * a.c2;
    s.k2 = kb + static_cast<size_t>(ii[1]) * a.c2;
    s.k3 = kb + static_cast<size_t>(ii[2]) * a.c2;
    s.w1 = ww[0];
    s.w2 = ww[1];
    s.w3 = ww[2];
    s.row = a.c1 ? a.skip + p * a.lds : nullptr;
  }
}

// element k of the logical A row (slow path: chunks that straddle a segment boundary / unaligned rows)
template <int PRO>
__device__ __forceinline__ float row_elem(const MlpArgs &a, const RowSrc &s, int k) {
  if (PRO == PRO_DENSE) return k < a.a_cols ? __ldg(s.row + k) : 0.f;
  if (PRO == PRO_SA_GATHER) {
    if (k < a.c_feat) return __ldg(s.row + k);
    const int d = k - a.c_feat;
    return d == 0 ? s.dx : d == 1 ? s.dy : d == 2 ? s.dz : 0.f;
  }
  if (k < a.c2)
    return __fmaf_rn(__ldg(s.k3 + k), s.w3, __fmaf_rn(__ldg(s.k1 + k), s.w1, __fmul_rn(__ldg(s.k2 + k), s.w2)));
  const int d = k - a.c2;
  return d < a.c1 ? __ldg(s.row + d) : 0.f;
}

// stage the 32 floats [k0, k0+32) of one row into the swizzled A tile
template <int PRO>
__device__ __forceinline__ void stage_a_row(const MlpArgs &a, const RowSrc &s, int r, int k0,
                                            uint32_t sa, bool vec_ok) {
  if (!s.live) {
#pragma unroll
    for (int c = 0; c < 8; ++c) sts128(sa + sw128_off(r, c), 0.f, 0.f, 0.f, 0.f);
    return;
  }
  // fast path: the whole 128-byte chunk comes from one 16-byte-aligned segment
  if (PRO == PRO_DENSE && vec_ok && k0 + 32 <= a.a_cols) {
    float4 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = ldg128(s.row + k0 + c * 4);
#pragma unroll
    for (int c = 0; c < 8; ++c)
      sts128(sa + sw128_off(r, c), to_tf32(v[c].x), to_tf32(v[c].y), to_tf32(v[c].z), to_tf32(v[c].w));
    return;
  }
  if (PRO == PRO_SA_GATHER && vec_ok && k0 + 32 <= a.c_feat) {
    float4 v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = ldg128(s.row + k0 + c * 4);
#pragma unroll
    for (int c = 0; c < 8; ++c)
      sts128(sa + sw128_off(r, c), to_tf32(v[c].x), to_tf32(v[c].y), to_tf32(v[c].z), to_tf32(v[c].w));
    return;
  }
  if (PRO == PRO_FP_INTERP && vec_ok && k0 + 32 <= a.c2) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 p1 = ldg128(s.k1 + k0 + c * 4), p2 = ldg128(s.k2 + k0 + c * 4),
                   p3 = ldg128(s.k3 + k0 + c * 4);
      // same contraction as three_interpolate (pn2_ops.cu): fma(p3,w3, fma(p1,w1, p2*w2))
      const float x = __fmaf_rn(p3.x, s.w3, __fmaf_rn(p1.x, s.w1, __fmul_rn(p2.x, s.w2)));
      const float y = __fmaf_rn(p3.y, s.w3, __fmaf_rn(p1.y, s.w1, __fmul_rn(p2.y, s.w2)));
      const float z = __fmaf_rn(p3.z, s.w3, __fmaf_rn(p1.z, s.w1, __fmul_rn(p2.z, s.w2)));
      const float w = __fmaf_rn(p3.w, s.w3, __fmaf_rn(p1.w, s.w1, __fmul_rn(p2.w, s.w2)));
      sts128(sa + sw128_off(r, c), to_tf32(x), to_tf32(y), to_tf32(z), to_tf32(w));
    }
    return;
  }
  if (PRO == PRO_FP_INTERP && vec_ok && k0 >= a.c2 && k0 - a.c2 + 32 <= a.c1 &&
      ((reinterpret_cast<uintptr_t>(s.row) & 15u) == 0)) {
    const float *src = s.row + (k0 - a.c2);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 v = ldg128(src + c * 4);
      sts128(sa + sw128_off(r, c), to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
    }
    return;
  }
  // generic path
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int k = k0 + c * 4;
    sts128(sa + sw128_off(r, c), to_tf32(row_elem<PRO>(a, s, k)), to_tf32(row_elem<PRO>(a, s, k + 1)),
           to_tf32(row_elem<PRO>(a, s, k + 2)), to_tf32(row_elem<PRO>(a, s, k + 3)));
  }
}

// ---- epilogue helpers ------------------------------------------------------------------------------
// After the call, lane l holds in v[0] the max over the 32 lanes of the ORIGINAL v[l] (column l).
__device__ __forceinline__ void warp_colmax_32(float (&v)[32], unsigned lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool hi = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float keep = hi ? v[i + o] : v[i];
      const float send = hi ? v[i] : v[i + o];
      v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, o));
    }
  }
}
// groups of 16 lanes: lane l (l' = l & 15) ends with columns 2l', 2l'+1 in v[0], v[1]
__device__ __forceinline__ void warp_colmax_16(float (&v)[32], unsigned lane) {
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) {
    const bool hi = (lane & o) != 0;
    const int half = o * 2;  // columns held after this step
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float keep = hi ? v[i + half] : v[i];
      const float send = hi ? v[i] : v[i + half];
      v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, o));
    }
  }
}
// groups of 8 lanes: lane l (l' = l & 7) ends with columns 4l'..4l'+3 in v[0..3]
__device__ __forceinline__ void warp_colmax_8(float (&v)[32], unsigned lane) {
#pragma unroll
  for (int o = 4; o >= 1; o >>= 1) {
    const bool hi = (lane & o) != 0;
    const int half = o * 4;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float keep = hi ? v[i + half] : v[i];
      const float send = hi ? v[i] : v[i + half];
      v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, o));
    }
  }
}

struct MlpSmemCtl {
  uint64_t full[kMlpMaxStages];
  uint64_t empty[kMlpMaxStages];
  uint64_t acc_full;
  uint32_t tmem_base;
};

template <int PRO, int EPI>
__global__ void __launch_bounds__(kMlpThreads, 1) mlp_layer_kernel(MlpArgs a) {
  extern __shared__ unsigned char mlp_smem_raw[];
  __shared__ MlpSmemCtl ctl;
  // 1024-byte aligned operand ring (SWIZZLE_128B atoms are 8 rows x 128 B)
  const uint32_t raw = smem_u32(mlp_smem_raw);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  const uint32_t a_bytes = kMlpBM * 128u;
  const uint32_t b_bytes = static_cast<uint32_t>(a.bn) * 128u;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023u) & ~1023u);

  const int t = threadIdx.x;
  const unsigned warp = t >> 5, lane = t & 31u;
  const long long p0 = static_cast<long long>(blockIdx.x) * kMlpBM;
  const int n0 = blockIdx.y * a.bn;
  const int bn = min(a.bn, a.n_pad - n0);  // this CTA's columns (multiple of 16)
  const int kc_total = a.k_pad / 32;
  const int S = a.stages;

  if (warp == 4) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(&ctl.full[s], 128);
        mbar_init(&ctl.empty[s], 1);
      }
      mbar_init(&ctl.acc_full, 1);
      mbar_fence_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&ctl.tmem_base)),
                 "r"(static_cast<uint32_t>(a.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctl.tmem_base;

  if (warp < 4) {
    // ================= producers: thread t stages row t of the A tile, all share the B tile =========
    RowSrc src;
    row_setup<PRO>(a, p0 + t, src);
    bool vec_ok = true;
    if (PRO == PRO_DENSE) vec_ok = (a.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.a) & 15u) == 0);
    if (PRO == PRO_SA_GATHER)
      vec_ok = a.c_feat > 0 && (a.ldf % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.feat) & 15u) == 0);
    if (PRO == PRO_FP_INTERP)
      vec_ok = (a.c2 % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.known_feat) & 15u) == 0);
    for (int kc = 0; kc < kc_total; ++kc) {
      const int s = kc % S;
      if (kc >= S) mbar_wait(&ctl.empty[s], static_cast<unsigned>((kc / S - 1) & 1));
      const uint32_t sa = ring + static_cast<uint32_t>(s) * stage_bytes;
      const uint32_t sb = sa + a_bytes;
      stage_a_row<PRO>(a, src, t, kc * 32, sa, vec_ok);
      // weights: rows n0..n0+bn of W, columns kc*32..+32 (already TF32-rounded and zero-padded)
      for (int i = t; i < bn * 8; i += 128) {
        const int n = i >> 3, c = i & 7;
        const float4 v = ldg128(a.w + static_cast<size_t>(n0 + n) * a.k_pad + kc * 32 + c * 4);
        sts128(sb + sw128_off(n, c), v.x, v.y, v.z, v.w);
      }
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor-core (async) proxy
      mbar_arrive(&ctl.full[s]);
    }

    // ================= epilogue: warp w owns TMEM lanes 32w..32w+31 = rows p0+32w.. ==================
    mbar_wait(&ctl.acc_full, 0);
    tc_fence_after();
    const long long prow = p0 + warp * 32 + lane;
    const uint32_t lane_addr = tmem + ((warp * 32u) << 16);
    for (int c0 = 0; c0 < bn; c0 += 32) {
      float v[32];
      const int cw = min(32, bn - c0);
      if (cw == 32) tmem_ld32(lane_addr + c0, v);
      else tmem_ld16(lane_addr + c0, v);
      if (EPI == EPI_STORE) {
        if (prow < a.rows) {
          float *o = a.out + prow * a.ldo + a.col0 + n0 + c0;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (q * 4 < cw) {
              float4 r;
              r.x = v[q * 4 + 0] + __ldg(a.bias + n0 + c0 + q * 4 + 0);
              r.y = v[q * 4 + 1] + __ldg(a.bias + n0 + c0 + q * 4 + 1);
              r.z = v[q * 4 + 2] + __ldg(a.bias + n0 + c0 + q * 4 + 2);
              r.w = v[q * 4 + 3] + __ldg(a.bias + n0 + c0 + q * 4 + 3);
              if (a.relu) {
                r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
              }
              *reinterpret_cast<float4 *>(o + q * 4) = r;
            }
          }
        }
      } else {
        // max over the `pool` rows of each centre; rows past the end contribute -inf
        if (prow >= a.rows) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = -__int_as_float(0x7f800000);
        }
        const long long grow0 = (p0 + warp * 32) / a.pool;  // first pooled row of this warp
        if (a.pool == 32) {
          warp_colmax_32(v, lane);
          const int col = static_cast<int>(lane);
          if (col < cw && p0 + warp * 32 < a.rows) {
            float r = v[0] + __ldg(a.bias + n0 + c0 + col);
            if (a.relu) r = fmaxf(r, 0.f);
            a.out[grow0 * a.ldo + a.col0 + n0 + c0 + col] = r;
          }
        } else if (a.pool == 16) {
          warp_colmax_16(v, lane);
          const int g = lane >> 4, col = static_cast<int>(lane & 15u) * 2;
          if (col < cw && p0 + warp * 32 + g * 16 < a.rows) {
            float2 r;
            r.x = v[0] + __ldg(a.bias + n0 + c0 + col);
            r.y = v[1] + __ldg(a.bias + n0 + c0 + col + 1);
            if (a.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); }
            *reinterpret_cast<float2 *>(a.out + (grow0 + g) * a.ldo + a.col0 + n0 + c0 + col) = r;
          }
        } else {  // pool == 8
          warp_colmax_8(v, lane);
          const int g = lane >> 3, col = static_cast<int>(lane & 7u) * 4;
          if (col < cw && p0 + warp * 32 + g * 8 < a.rows) {
            float4 r;
            r.x = v[0] + __ldg(a.bias + n0 + c0 + col);
            r.y = v[1] + __ldg(a.bias + n0 + c0 + col + 1);
            r.z = v[2] + __ldg(a.bias + n0 + c0 + col + 2);
            r.w = v[3] + __ldg(a.bias + n0 + c0 + col + 3);
            if (a.relu) {
              r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
            }
            *reinterpret_cast<float4 *>(a.out + (grow0 + g) * a.ldo + a.col0 + n0 + c0 + col) = r;
          }
        }
      }
    }
    tc_fence_before();
  } else {
    // ================= warp 4: MMA issuer =============================================================
    const uint32_t idesc = instr_desc_tf32(bn);
    for (int kc = 0; kc < kc_total; ++kc) {
      const int s = kc % S;
      mbar_wait(&ctl.full[s], static_cast<unsigned>((kc / S) & 1));
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = ring + static_cast<uint32_t>(s) * stage_bytes;
        const uint64_t adesc = smem_desc_sw128(sa), bdesc = smem_desc_sw128(sa + a_bytes);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)  // K = 8 tf32 = 32 bytes per instruction: +2 in 16-byte units
          umma_tf32(tmem, adesc + static_cast<uint64_t>(k4 * 2), bdesc + static_cast<uint64_t>(k4 * 2),
                    idesc, (kc > 0 || k4 > 0) ? 1u : 0u);
        umma_commit(&ctl.empty[s]);  // stage reusable once these MMAs have read it
        if (kc == kc_total - 1) umma_commit(&ctl.acc_full);
      }
      __syncwarp();
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                 "r"(static_cast<uint32_t>(a.tmem_cols))
                 : "memory");
  }
}

template <int PRO, int EPI>
int launch_mlp(MlpArgs &a, cudaStream_t st) {
  if (a.rows <= 0) return PVN3D_OK;
  if (a.k_pad <= 0 || a.k_pad % 32 || a.n_pad <= 0 || a.n_pad % 16) return PVN3D_ERR_INVALID_ARG;
  // columns per CTA: <= 256, multiple of 16, as even a split as possible
  const int nblk = ceil_div(a.n_pad, 256);
  a.bn = ((ceil_div(a.n_pad, nblk) + 15) / 16) * 16;
  int tc = 32;
  while (tc < a.bn) tc <<= 1;
  a.tmem_cols = tc;
  const size_t stage_bytes = kMlpBM * 128 + align_up(static_cast<size_t>(a.bn) * 128, 1024);
  int stages = static_cast<int>((200 * 1024) / stage_bytes);
  if (stages > kMlpMaxStages) stages = kMlpMaxStages;
  if (stages > a.k_pad / 32) stages = a.k_pad / 32;
  if (stages < 1) stages = 1;
  a.stages = stages;
  const size_t smem = stages * stage_bytes + 1024;
  auto kern = mlp_layer_kernel<PRO, EPI>;
  static PerDeviceOnce once;
  if (once.first_time())
    PVN3D_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024),
                   "mlp smem attr");
  const long long tiles = (a.rows + kMlpBM - 1) / kMlpBM;
  if (tiles > 0x7fffffffll) return PVN3D_ERR_UNSUPPORTED;
  dim3 grid(static_cast<unsigned>(tiles), ceil_div(a.n_pad, a.bn));
  kern<<<grid, kMlpThreads, smem, st>>>(a);
  return check_launch("mlp_layer_kernel");
}

int dispatch(MlpArgs &a, int pro, int pool, cudaStream_t st) {
  if (pool) {
    if (pool != 8 && pool != 16 && pool != 32) return PVN3D_ERR_UNSUPPORTED;
    a.pool = pool;
    if (pro == PRO_DENSE) return launch_mlp<PRO_DENSE, EPI_MAXPOOL>(a, st);
    if (pro == PRO_SA_GATHER) return launch_mlp<PRO_SA_GATHER, EPI_MAXPOOL>(a, st);
    return PVN3D_ERR_INVALID_ARG;
  }
  a.pool = 0;
  if (pro == PRO_DENSE) return launch_mlp<PRO_DENSE, EPI_STORE>(a, st);
  if (pro == PRO_SA_GATHER) return launch_mlp<PRO_SA_GATHER, EPI_STORE>(a, st);
  return launch_mlp<PRO_FP_INTERP, EPI_STORE>(a, st);
}

// inverse-distance weights of PointnetFPModule.forward (pointnet2_modules.py:183-186) + 3-NN indices
__global__ void nn_weights_kernel(const float *__restrict__ dist2, long long rows,
                                  float *__restrict__ w) {
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= rows) return;
  const float r1 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(dist2[p * 3 + 0]), 1e-8f));
  const float r2 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(dist2[p * 3 + 1]), 1e-8f));
  const float r3 = __fdiv_rn(1.0f, __fadd_rn(__fsqrt_rn(dist2[p * 3 + 2]), 1e-8f));
  const float norm = __fadd_rn(__fadd_rn(r1, r2), r3);
  w[p * 3 + 0] = __fdiv_rn(r1, norm);
  w[p * 3 + 1] = __fdiv_rn(r2, norm);
  w[p * 3 + 2] = __fdiv_rn(r3, norm);
}

}  // namespace
}  // namespace pvn3d

using namespace pvn3d;

extern "C" int pvn3d_mlp_dense(const float *a, int lda, int a_cols, long long rows, const float *w,
                               const float *bias, int k_pad, int n_pad, int relu, int pool,
                               float *out, int ldo, int col0, pvn3d_stream_t stream) {
  if (!a || !w || !bias || !out || lda < a_cols || a_cols < 0 || a_cols % 4 || rows < 0 || ldo % 4 ||
      col0 % 4)
    return PVN3D_ERR_INVALID_ARG;
  if (pool && rows % pool) return PVN3D_ERR_INVALID_ARG;
  MlpArgs m{};
  m.w = w; m.bias = bias; m.rows = rows; m.k_pad = k_pad; m.n_pad = n_pad;
  m.a = a; m.lda = lda; m.a_cols = a_cols;
  m.out = out; m.ldo = ldo; m.col0 = col0; m.relu = relu;
  return dispatch(m, PRO_DENSE, pool, as_stream(stream));
}

extern "C" int pvn3d_mlp_sa_first(const float *xyz, const float *new_xyz, const float *feat_pm,
                                  int ldf, int c_feat, const int *idx, int b, int n, int m, int ns,
                                  const float *w, const float *bias, int k_pad, int n_pad, int relu,
                                  int pool, float *out, int ldo, int col0, pvn3d_stream_t stream) {
  if (!xyz || !new_xyz || !idx || !w || !bias || !out || b < 0 || n <= 0 || m < 0 || ns <= 0 ||
      c_feat < 0 || (c_feat > 0 && (!feat_pm || ldf < c_feat)) || k_pad < c_feat + 3 || ldo % 4 ||
      col0 % 4)
    return PVN3D_ERR_INVALID_ARG;
  if (pool && pool != ns) return PVN3D_ERR_INVALID_ARG;
  MlpArgs a{};
  a.w = w; a.bias = bias; a.rows = static_cast<long long>(b) * m * ns; a.k_pad = k_pad; a.n_pad = n_pad;
  a.xyz = xyz; a.new_xyz = new_xyz; a.feat = feat_pm; a.ldf = ldf; a.c_feat = c_feat; a.idx = idx;
  a.n = n; a.m = m; a.ns = ns;
  a.out = out; a.ldo = ldo; a.col0 = col0; a.relu = relu;
  return dispatch(a, PRO_SA_GATHER, pool, as_stream(stream));
}

extern "C" int pvn3d_mlp_fp_first(const float *known_feat_pm, int c2, const int *nn_idx,
                                  const float *nn_w, const float *skip_pm, int lds, int c1, int b,
                                  int n_unknown, int m_known, const float *w, const float *bias,
                                  int k_pad, int n_pad, int relu, float *out, int ldo, int col0,
                                  pvn3d_stream_t stream) {
  if (!known_feat_pm || !nn_idx || !nn_w || !w || !bias || !out || b < 0 || n_unknown < 0 ||
      m_known <= 0 || c2 <= 0 || c1 < 0 || (c1 > 0 && (!skip_pm || lds < c1)) || k_pad < c2 + c1 ||
      ldo % 4 || col0 % 4)
    return PVN3D_ERR_INVALID_ARG;
  MlpArgs a{};
  a.w = w; a.bias = bias; a.rows = static_cast<long long>(b) * n_unknown; a.k_pad = k_pad; a.n_pad = n_pad;
  a.known_feat = known_feat_pm; a.c2 = c2; a.nn_idx = nn_idx; a.nn_w = nn_w; a.skip = skip_pm;
  a.lds = lds; a.c1 = c1; a.n_unknown = n_unknown; a.m_known = m_known;
  a.out = out; a.ldo = ldo; a.col0 = col0; a.relu = relu;
  return dispatch(a, PRO_FP_INTERP, 0, as_stream(stream));
}

extern "C" int pvn3d_three_nn_weights(const float *dist2, long long rows, float *weight,
                                      pvn3d_stream_t stream) {
  if (!dist2 || !weight || rows < 0) return PVN3D_ERR_INVALID_ARG;
  if (rows == 0) return PVN3D_OK;
  nn_weights_kernel<<<static_cast<unsigned>((rows + 255) / 256), 256, 0, as_stream(stream)>>>(
      dist2, rows, weight);
  return check_launch("nn_weights_kernel");
}
