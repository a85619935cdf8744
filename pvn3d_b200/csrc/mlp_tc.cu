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
// A-operand producers (warps 4-11, eight lanes per row of the 128-row tile):
//   DENSE      rows of a point-major activation matrix [P, lda]
//   SA_GATHER  row (b,j,s): [ feat_pm[b, idx[b,j,s], 0:C] | xyz[b,idx] - new_xyz[b,j] | 0 ... ]
//              (QueryAndGroup.forward, pointnet2_utils.py:311-321, never materialised; the weight
//              matrix has its xyz columns moved behind the feature columns to keep rows 16-B aligned)
//   FP_INTERP  row (b,j): [ sum_t w_t * known_feat_pm[b, idx_t, 0:C2] | skip[b, j, 0:C1] | 0 ... ]
//              (three_interpolate + torch.cat of PointnetFPModule.forward, pointnet2_modules.py:188-199)
// Operands are staged in shared memory in the canonical K-major SWIZZLE_128B layout (32 fp32 = 128 B
// per row per stage; weights and pre-rounded activations arrive by cp.async), one elected thread of
// warp 12 issues tcgen05.mma (M=128, N<=256, K=8 per instruction) into one of two TMEM accumulators
// and releases stages with tcgen05.commit; the kernel is persistent (one CTA per SM, tiles strided),
// so staging of tile j+1 and the epilogue of tile j-1 overlap the MMAs of tile j.  The epilogue warps
// (0-3) read the accumulator with
// tcgen05.ld (32x32b), add the folded-BN bias, apply ReLU and either store the point-major row or
// reduce over the nsample rows of each centre with a transposing shuffle butterfly (max-pool) and
// store one 128-byte line per centre.
//
// TF32: operands are rounded to nearest (cvt.rna.tf32.f32) when staged, accumulation is fp32 -- the
// precision class of the reference's default cuDNN convolutions (torch.backends.cudnn.allow_tf32).
#include <cuda.h>  // CUtensorMap + the cuTensorMapEncodeTiled prototype (entry point fetched through the runtime)

#include "common.cuh"

namespace pvn3d {
namespace {

constexpr int kMlpBM = 128;
constexpr int kMlpEpiWarps = 4;  // warps 0-3: epilogue (warp w owns TMEM lanes 32w..32w+31)
constexpr int kMlpProWarps = 8;  // warps 4-11: two producer groups of 128 threads, alternate K chunks
constexpr int kMlpThreads = (kMlpEpiWarps + kMlpProWarps + 1) * 32;  // warp 12: TMEM owner + MMA issuer
constexpr int kMlpMaxStages = 6;

enum : int { PRO_DENSE = 0, PRO_SA_GATHER = 1, PRO_FP_INTERP = 2, PRO_SA_FACT = 3, PRO_FP_FACT = 4 };
enum : int { EPI_STORE = 0, EPI_MAXPOOL = 1, EPI_SUMPOOL = 2, EPI_MAXPOOL_T = 3, EPI_STORE_T = 4 };

struct MlpArgs {
  // W as a TMA tensor map ([n_pad][k_pad] fp32, box 32 columns x bn rows, SWIZZLE_128B): one
  // cp.async.bulk.tensor.2d per K chunk lands the B operand in the UMMA layout (use_tma, else cp.async)
  alignas(64) CUtensorMap tmap;
  int use_tma;
  // GEMM
  const float *w;     // [n_pad][k_pad]
  const float *bias;  // [n_pad]
  long long rows;     // P
  int k_pad;          // multiple of 32
  int n_pad;          // multiple of 16
  int bn;             // columns per tile (multiple of 16, <= 256)
  int stages;
  int tmem_cols;      // power of two >= max(32, bn); acc_bufs accumulators are allocated
  int acc_bufs;       // 2: the epilogue of tile j overlaps the MMAs of tile j+1; 1: wide tiles (tmem_cols = 256) of two
                      // co-resident CTAs -- the other CTA's MMAs fill the gap
  // DENSE
  const float *a;
  int lda;
  int a_cols;  // valid columns of a (multiple of 4); the rest of k_pad reads as zero
  int a_tf32;  // a is already TF32-rounded and 16-byte aligned: copied with cp.async, no registers
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
  int round_out;  // store TF32-rounded values (the next layer then takes them with a_tf32)
  int pool;       // nsample of the max-pool epilogue (8, 16 or 32)
  int reserve_sms;  // host only: SMs left to concurrent kernels (PVN3D_MLP_RESERVE_SMS in flags)
  int bias_npb;     // > 0: bias is [rows / bias_npb][n_pad] -- one vector per batch element (bias_npb % 128 == 0)
  int out_cn;       // STORE_T: points per frame; out is [rows / out_cn][n_pad][out_cn] (channel-major frames, out_cn % 32 == 0)
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
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 ldg128(const float *p) {
  return __ldg(reinterpret_cast<const float4 *>(p));
}
// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only); src_bytes = 0 writes zeros
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, unsigned src_bytes = 16u) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// register transaction bytes on an mbarrier WITHOUT arriving (the thread arrives later like every producer)
__device__ __forceinline__ void mbar_expect_tx_only(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// one 2-D tile global -> shared through a tensor map (TMA), completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap *map, int x, int y,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
          "r"(dst_smem),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
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

// ---- A-operand producers ---------------------------------------------------------------------------
// A producer group is 128 threads = 4 warps; warp pw stages rows 32pw..32pw+31 of the 128-row tile.
// Inside a warp, 8 consecutive lanes share a row (lane & 7 = the 16-byte chunk of the 128-byte K
// slice), so one warp-wide LDG.128 reads four full 128-byte lines and one STS.128 fills four swizzled
// rows without bank conflicts; the 32 rows take 8 passes (row of pass j = 32pw + 4j + lane/8), whose 8
// loads are all in flight before the first store.
struct RowState {
  unsigned live;        // bit j: row of pass j exists (p < rows)
  const float *row[8];  // DENSE: a + p*lda;  SA: feature row of the grouped point
  int qrow[8], crow[8];        // SA: row of the grouped point in xyz (b*n + idx), of its centre (b*m + j)
  int g1[8], g2[8], g3[8];     // FP: rows of the three neighbours in known_feat (b*m_known + idx)
  float w1[8], w2[8], w3[8];   // FP: their weights
};

// rows p_first, p_first+4, ..: batch index and offset inside the batch with ONE 64-bit division
template <int PRO>
__device__ __forceinline__ void rows_setup(const MlpArgs &a, long long p_first, RowState &s) {
  s.live = 0u;
  const unsigned per_b = (PRO == PRO_SA_GATHER || PRO == PRO_SA_FACT)
                             ? static_cast<unsigned>(a.m) * static_cast<unsigned>(a.ns)
                             : static_cast<unsigned>(a.n_unknown);
  unsigned b0 = 0, rem0 = 0;
  if (PRO != PRO_DENSE) {
    const long long pf = p_first < a.rows ? p_first : 0;
    b0 = static_cast<unsigned>(pf / per_b);
    rem0 = static_cast<unsigned>(pf - static_cast<long long>(b0) * per_b);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const long long p = p_first + 4 * j;
    const bool live = p < a.rows;
    if (live) s.live |= 1u << j;
    const long long pc = live ? p : 0;
    unsigned rem = rem0 + 4u * j, b = b0;
    if (PRO != PRO_DENSE) {
      if (rem >= per_b) {
        const unsigned over = rem / per_b;
        b += over;
        rem -= over * per_b;
      }
      if (!live) { b = 0; rem = 0; }
    }
    if (PRO == PRO_DENSE) {
      s.row[j] = a.a + pc * a.lda;
    } else if (PRO == PRO_SA_GATHER || PRO == PRO_SA_FACT) {
      const int q = __ldg(a.idx + pc);
      s.qrow[j] = static_cast<int>(b) * a.n + q;
      s.crow[j] = static_cast<int>(b) * a.m + static_cast<int>(rem / static_cast<unsigned>(a.ns));
      s.row[j] = a.feat + static_cast<size_t>(s.qrow[j]) * a.ldf;  // never read when c_feat == 0
    } else {
      const int base = static_cast<int>(b) * a.m_known;
      s.g1[j] = base + __ldg(a.nn_idx + pc * 3 + 0);
      s.g2[j] = base + __ldg(a.nn_idx + pc * 3 + 1);
      s.g3[j] = base + __ldg(a.nn_idx + pc * 3 + 2);
      s.w1[j] = __ldg(a.nn_w + pc * 3 + 0);
      s.w2[j] = __ldg(a.nn_w + pc * 3 + 1);
      s.w3[j] = __ldg(a.nn_w + pc * 3 + 2);
    }
  }
}

__device__ __forceinline__ void sts_tf32(uint32_t addr, const float4 &v) {
  sts128(addr, to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
}

// stage columns [k0 + 4*sub, +4) of this thread's 8 rows into the swizzled A tile at `sa`
template <int PRO>
__device__ __forceinline__ void stage_a_chunk(const MlpArgs &a, const RowState &s, long long p_first,
                                              int r_first, int sub, int k0, uint32_t sa, bool vec_ok) {
  const int k = k0 + 4 * sub;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  if (PRO == PRO_FP_FACT) {
    // second layer of a FACTORED FP module: row (b,j) = relu( sum_t w_t * P[b, idx_t, :] + S[b, j, :] ) with
    // P = W1k . known (once per KNOWN point) and S = W1s . skip + b1 (the skip columns only): the first layer is
    // linear before its ReLU, so three_interpolate commutes with it (pointnet2_modules.py:183-204)
    if (k >= a.c2) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sts128(sa + sw128_off(r_first + 4 * j, sub), 0.f, 0.f, 0.f, 0.f);
      return;
    }
#pragma unroll
    for (int h = 0; h < 4; ++h) {   // four quarters: 8 LDG.128 in flight each
      float4 p1[2], p2[2], p3[2], sk[2];
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = h * 2 + jj;
        const float *pf = a.known_feat + k;
        p1[jj] = ldg128(pf + static_cast<size_t>(s.g1[j]) * a.c2);
        p2[jj] = ldg128(pf + static_cast<size_t>(s.g2[j]) * a.c2);
        p3[jj] = ldg128(pf + static_cast<size_t>(s.g3[j]) * a.c2);
        const long long pr = ((s.live >> j) & 1u) ? p_first + 4 * j : 0;
        sk[jj] = ldg128(a.skip + pr * a.c2 + k);
      }
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = h * 2 + jj;
        float4 v;
        v.x = fmaxf(__fmaf_rn(p3[jj].x, s.w3[j], __fmaf_rn(p1[jj].x, s.w1[j], __fmul_rn(p2[jj].x, s.w2[j]))) + sk[jj].x, 0.f);
        v.y = fmaxf(__fmaf_rn(p3[jj].y, s.w3[j], __fmaf_rn(p1[jj].y, s.w1[j], __fmul_rn(p2[jj].y, s.w2[j]))) + sk[jj].y, 0.f);
        v.z = fmaxf(__fmaf_rn(p3[jj].z, s.w3[j], __fmaf_rn(p1[jj].z, s.w1[j], __fmul_rn(p2[jj].z, s.w2[j]))) + sk[jj].z, 0.f);
        v.w = fmaxf(__fmaf_rn(p3[jj].w, s.w3[j], __fmaf_rn(p1[jj].w, s.w1[j], __fmul_rn(p2[jj].w, s.w2[j]))) + sk[jj].w, 0.f);
        if (!((s.live >> j) & 1u)) v = zero;
        sts_tf32(sa + sw128_off(r_first + 4 * j, sub), v);
      }
    }
    return;
  }
  if (PRO == PRO_SA_FACT) {
    // second layer of a FACTORED SA scale: row (b,i,s) = relu(U[b, idx[b,i,s], :] - V[b, i, :]), where
    // U = W1 . [f_j | x_j] for every POINT and V = W1x . c_i - bias1 for every CENTRE (the first layer is linear
    // before its ReLU, so it is evaluated once per point instead of once per (centre, neighbour) pair)
    if (k >= a.c_feat) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sts128(sa + sw128_off(r_first + 4 * j, sub), 0.f, 0.f, 0.f, 0.f);
      return;
    }
    // the thread's 8 rows are 4 apart inside one 32-row block: at most two centres (nsample 16), one for nsample 32:
    // two V loads + all eight U loads are in flight together
    const float4 va = ldg128(a.new_xyz + static_cast<size_t>(s.crow[0]) * a.ldf + k);
    const float4 vb = ldg128(a.new_xyz + static_cast<size_t>(s.crow[4]) * a.ldf + k);
    float4 u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] = ldg128(s.row[j] + k);
    const bool two = a.ns % 16 == 0;   // 16-row groups never straddle centres: rows j < 4 share crow[0], j >= 4 crow[4]
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 v = j < 4 ? va : vb;
      if (!two) v = ldg128(a.new_xyz + static_cast<size_t>(s.crow[j]) * a.ldf + k);   // other nsample: per-row centre
      float4 r = make_float4(fmaxf(u[j].x - v.x, 0.f), fmaxf(u[j].y - v.y, 0.f), fmaxf(u[j].z - v.z, 0.f),
                             fmaxf(u[j].w - v.w, 0.f));
      if (!((s.live >> j) & 1u)) r = zero;
      sts_tf32(sa + sw128_off(r_first + 4 * j, sub), r);
    }
    return;
  }
  if (PRO == PRO_DENSE || PRO == PRO_SA_GATHER) {
    const int seg = PRO == PRO_DENSE ? a.a_cols : a.c_feat;        // vector-loadable prefix of the row
    const int end = PRO == PRO_DENSE ? a.a_cols : a.c_feat + 3;    // logical row length
    if (vec_ok && k + 4 <= seg) {
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (s.live >> j) & 1u ? ldg128(s.row[j] + k) : zero;
#pragma unroll
      for (int j = 0; j < 8; ++j) sts_tf32(sa + sw128_off(r_first + 4 * j, sub), v[j]);
      return;
    }
    if (k >= end) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sts128(sa + sw128_off(r_first + 4 * j, sub), 0.f, 0.f, 0.f, 0.f);
      return;
    }
  } else {
    if (vec_ok && k + 4 <= a.c2) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {  // two halves: 12 LDG.128 in flight each
        float4 p1[4], p2[4], p3[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = h * 4 + jj;
          const float *kf = a.known_feat + k;
          p1[jj] = ldg128(kf + static_cast<size_t>(s.g1[j]) * a.c2);
          p2[jj] = ldg128(kf + static_cast<size_t>(s.g2[j]) * a.c2);
          p3[jj] = ldg128(kf + static_cast<size_t>(s.g3[j]) * a.c2);
        }
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = h * 4 + jj;
          // same contraction as three_interpolate (pn2_ops.cu): fma(p3,w3, fma(p1,w1, p2*w2))
          float4 v;
          v.x = __fmaf_rn(p3[jj].x, s.w3[j], __fmaf_rn(p1[jj].x, s.w1[j], __fmul_rn(p2[jj].x, s.w2[j])));
          v.y = __fmaf_rn(p3[jj].y, s.w3[j], __fmaf_rn(p1[jj].y, s.w1[j], __fmul_rn(p2[jj].y, s.w2[j])));
          v.z = __fmaf_rn(p3[jj].z, s.w3[j], __fmaf_rn(p1[jj].z, s.w1[j], __fmul_rn(p2[jj].z, s.w2[j])));
          v.w = __fmaf_rn(p3[jj].w, s.w3[j], __fmaf_rn(p1[jj].w, s.w1[j], __fmul_rn(p2[jj].w, s.w2[j])));
          if (!((s.live >> j) & 1u)) v = zero;
          sts_tf32(sa + sw128_off(r_first + 4 * j, sub), v);
        }
      }
      return;
    }
    const bool skip_vec = (a.c2 % 4 == 0) && (a.lds % 4 == 0) &&
                          ((reinterpret_cast<uintptr_t>(a.skip) & 15u) == 0);
    if (skip_vec && k >= a.c2 && k - a.c2 + 4 <= a.c1) {
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        v[j] = (s.live >> j) & 1u ? ldg128(a.skip + (p_first + 4 * j) * a.lds + (k - a.c2)) : zero;
#pragma unroll
      for (int j = 0; j < 8; ++j) sts_tf32(sa + sw128_off(r_first + 4 * j, sub), v[j]);
      return;
    }
    if (k >= a.c2 + a.c1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) sts128(sa + sw128_off(r_first + 4 * j, sub), 0.f, 0.f, 0.f, 0.f);
      return;
    }
  }
  // generic path (chunks that straddle a segment boundary / unaligned rows).  Which source an element
  // comes from depends only on its column, not on the row, so the loads of the 8 passes are issued
  // together as predicated loads -- no per-element branches, one round trip instead of 32.
  float vv[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const bool live = (s.live >> j) & 1u;
    const long long p = p_first + 4 * j;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int kk = k + e;
      float v = 0.f;
      if (PRO == PRO_DENSE) {
        const bool on = live && kk < a.a_cols;
        v = on ? __ldg(s.row[j] + (on ? kk : 0)) : 0.f;
      } else if (PRO == PRO_SA_GATHER) {
        const bool isf = live && kk < a.c_feat;
        const int d = kk - a.c_feat;
        const bool isx = live && d >= 0 && d <= 2;
        const float f = isf ? __ldg(s.row[j] + (isf ? kk : 0)) : 0.f;
        // grouped_xyz -= new_xyz (pointnet2_utils.py:314)
        const float px = isx ? __ldg(a.xyz + static_cast<size_t>(s.qrow[j]) * 3 + (isx ? d : 0)) : 0.f;
        const float pc = isx ? __ldg(a.new_xyz + static_cast<size_t>(s.crow[j]) * 3 + (isx ? d : 0)) : 0.f;
        v = isf ? f : px - pc;
      } else {
        const bool isk = live && kk < a.c2;
        const int d = kk - a.c2;
        const bool iss = live && d >= 0 && d < a.c1;
        const float *kf = a.known_feat + (isk ? kk : 0);
        const float p1 = isk ? __ldg(kf + static_cast<size_t>(s.g1[j]) * a.c2) : 0.f;
        const float p2 = isk ? __ldg(kf + static_cast<size_t>(s.g2[j]) * a.c2) : 0.f;
        const float p3 = isk ? __ldg(kf + static_cast<size_t>(s.g3[j]) * a.c2) : 0.f;
        const float sk = iss ? __ldg(a.skip + p * a.lds + (iss ? d : 0)) : 0.f;
        // same contraction as three_interpolate (pn2_ops.cu): fma(p3,w3, fma(p1,w1, p2*w2))
        v = isk ? __fmaf_rn(p3, s.w3[j], __fmaf_rn(p1, s.w1[j], __fmul_rn(p2, s.w2[j]))) : sk;
      }
      vv[j][e] = v;
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    sts_tf32(sa + sw128_off(r_first + 4 * j, sub), make_float4(vv[j][0], vv[j][1], vv[j][2], vv[j][3]));
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
// After the call, lane l holds in v[0] the SUM over the 32 lanes of the ORIGINAL v[l] (fixed order: reproducible).
__device__ __forceinline__ void warp_colsum_32(float (&v)[32], unsigned lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool hi = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float keep = hi ? v[i + o] : v[i];
      const float send = hi ? v[i] : v[i + o];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
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
  uint64_t full[kMlpMaxStages];   // 128 arrivals: every producer thread of the filling group
  uint64_t empty[kMlpMaxStages];  // 1 arrival: tcgen05.commit of the MMAs that read the stage
  uint64_t acc_full[2];           // 1 arrival: tcgen05.commit after a tile's last MMA
  uint64_t acc_empty[2];          // 128 arrivals: the epilogue threads, once they have read it
  uint32_t tmem_base;
};

// Persistent, warp-specialised: CTA c works on tiles c, c+grid, ... (tile = 128 rows x bn columns).
//   producers  fill a ring of K-chunk stages (A: 128 rows x 128 B, B: bn rows x 128 B), running ahead
//              of the tensor core across tile boundaries;
//   warp 12    issues the MMAs of tile j into accumulator j&1 of TMEM;
//   epilogue   drains accumulator j&1 while the MMAs of tile j+1 fill the other one.
// OCC = CTAs per SM the kernel is compiled for.  The layers are latency-bound (gathers through L2, TMEM
// round trips) at 13 warps per SM; two co-resident CTAs (<= 78 registers per thread, half the operand ring,
// two accumulators of <= 128 TMEM columns each) double the loads in flight for the narrow layers.
template <int PRO, int EPI, int OCC>
__global__ void __launch_bounds__(kMlpThreads, OCC) mlp_layer_kernel(const __grid_constant__ MlpArgs a) {
  // dynamic shared memory only, 1024-byte aligned: [operand ring: stages x stage_bytes][epilogue staging 4 x 4 KB]
  // [barriers].  No static block and no alignment slack: two CTAs with 96 KB rings each fit one SM's 228 KB.
  extern __shared__ __align__(1024) unsigned char mlp_smem_al[];
  const uint32_t ring = smem_u32(mlp_smem_al);   // SWIZZLE_128B atoms are 8 rows x 128 B
  if (ring & 1023u) __trap();
  const uint32_t a_bytes = kMlpBM * 128u;
  const uint32_t stage_bytes = a_bytes + ((static_cast<uint32_t>(a.bn) * 128u + 1023u) & ~1023u);
  MlpSmemCtl &ctl = *reinterpret_cast<MlpSmemCtl *>(mlp_smem_al + static_cast<size_t>(a.stages) * stage_bytes +
                                                    kMlpEpiWarps * 4096);
  const unsigned nbuf = static_cast<unsigned>(a.acc_bufs);

  const int t = threadIdx.x;
  const unsigned warp = t >> 5, lane = t & 31u;
  const int kc_total = a.k_pad / 32;
  const int S = a.stages;
  const long long row_tiles = (a.rows + kMlpBM - 1) / kMlpBM;
  const int n_blocks = (a.n_pad + a.bn - 1) / a.bn;
  const long long total_tiles = row_tiles * n_blocks;

  if (warp == kMlpEpiWarps + kMlpProWarps) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(&ctl.full[s], 128);
        mbar_init(&ctl.empty[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&ctl.acc_full[b], 1);
        mbar_init(&ctl.acc_empty[b], 128);
      }
      mbar_fence_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&ctl.tmem_base)),
                 "r"(nbuf * static_cast<uint32_t>(a.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctl.tmem_base;

  if (warp >= kMlpEpiWarps && warp < kMlpEpiWarps + kMlpProWarps) {
    // ================= producers ======================================================================
    const int pt = (t - kMlpEpiWarps * 32) & 127;          // thread inside the group
    const unsigned grp = (warp - kMlpEpiWarps) >> 2;       // group 0 / 1 takes even / odd chunks
    const int pw = pt >> 5, sub = static_cast<int>(lane & 7u), rg = static_cast<int>(lane >> 3);
    const int r_first = 32 * pw + rg;
    bool vec_ok = true;
    if (PRO == PRO_DENSE) vec_ok = (a.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.a) & 15u) == 0);
    if (PRO == PRO_SA_GATHER)
      vec_ok = a.c_feat > 0 && (a.ldf % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.feat) & 15u) == 0);
    if (PRO == PRO_FP_INTERP)
      vec_ok = (a.c2 % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.known_feat) & 15u) == 0);
    // pre-rounded rows are copied global -> shared asynchronously: DENSE activations of a ROUND_OUT layer, and
    // the descriptor columns of SA_GATHER rows when the level table was stored rounded (PVN3D_MLP_A_TF32);
    // the xyz / padding chunk of a gathered row and everything else is staged through registers
    const bool a_async = (PRO == PRO_DENSE || PRO == PRO_SA_GATHER) && a.a_tf32 && vec_ok;
    const int async_cols = PRO == PRO_DENSE ? a.k_pad : (a.c_feat / 32) * 32;   // chunks [0, async_cols) are asynchronous
    int pend0 = 0, pend1 = 0, npend = 0;  // stages whose copies are committed but not yet published
    long long it_base = 0;  // number of K chunks staged before this tile (same in every role)
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it_base += kc_total) {
      if (kc_total == 1 && static_cast<unsigned>(it_base & 1) != grp) continue;  // other group's tile
      const long long rt = tile / n_blocks;
      const int nb = static_cast<int>(tile - rt * n_blocks);
      const long long p0 = rt * kMlpBM;
      const int n0 = nb * a.bn;
      const int bn = min(a.bn, a.n_pad - n0);
      const long long p_first = p0 + r_first;
      RowState rs;
      rows_setup<PRO>(a, p_first, rs);
      for (int kc = 0; kc < kc_total; ++kc) {
        const long long it = it_base + kc;
        if (static_cast<unsigned>(it & 1) != grp) continue;
        const int s = static_cast<int>(it % S);
        // asynchronous path: two chunks of copies stay in flight per thread; the older one is
        // published (complete -> proxy fence -> arrive) before this thread can block on a free stage
        if (npend == 2) {
          cp_async_wait<1>();
          fence_proxy_async_smem();  // landed cp.async data (generic proxy) -> tensor-core (async) proxy
          mbar_arrive(&ctl.full[pend0]);
          pend0 = pend1;
          npend = 1;
        }
        mbar_wait(&ctl.empty[s], static_cast<unsigned>(((it / S) & 1) ^ 1));
        const uint32_t sa = ring + static_cast<uint32_t>(s) * stage_bytes;
        const uint32_t sb = sa + a_bytes;
        // weights: rows n0..n0+bn of W, columns kc*32..+32 (TF32-rounded, zero-padded): global -> smem
        if (a.use_tma) {
          if (pt == 0) {   // one tensor-map copy, completing on the stage's `full` barrier in bytes
            mbar_expect_tx_only(&ctl.full[s], static_cast<unsigned>(a.bn) * 128u);
            tma_load_2d(sb, &a.tmap, kc * 32, n0, &ctl.full[s]);
          }
        } else {
          for (int i = pt; i < bn * 8; i += 128) {
            const int n = i >> 3, c = i & 7;
            cp_async16(sb + sw128_off(n, c), a.w + static_cast<size_t>(n0 + n) * a.k_pad + kc * 32 + c * 4);
          }
        }
        if (a_async && kc * 32 < async_cols) {
          const int k = kc * 32 + 4 * sub;
          const int valid = PRO == PRO_DENSE ? a.a_cols : a.c_feat;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            cp_async16(sa + sw128_off(r_first + 4 * j, sub), rs.row[j] + (k < valid ? k : 0),
                       (((rs.live >> j) & 1u) && k < valid) ? 16u : 0u);
          cp_async_commit();
          if (npend == 0) pend0 = s; else pend1 = s;
          ++npend;
        } else {
          cp_async_commit();
          stage_a_chunk<PRO>(a, rs, p_first, r_first, sub, kc * 32, sa, vec_ok);
          cp_async_wait<0>();        // the weights of this chunk and every asynchronous chunk still pending have landed
          fence_proxy_async_smem();  // generic-proxy stores / copies -> visible to the tensor-core proxy
          if (npend >= 1) mbar_arrive(&ctl.full[pend0]);
          if (npend == 2) mbar_arrive(&ctl.full[pend1]);
          npend = 0;
          mbar_arrive(&ctl.full[s]);
        }
      }
    }
    if (npend == 2) {
      cp_async_wait<1>();
      fence_proxy_async_smem();
      mbar_arrive(&ctl.full[pend0]);
      pend0 = pend1;
      npend = 1;
    }
    if (npend == 1) {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      mbar_arrive(&ctl.full[pend0]);
    }
  } else if (warp < kMlpEpiWarps) {
    // ================= epilogue: warp w owns TMEM lanes 32w..32w+31 = rows p0+32w.. ==================
    long long j = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++j) {
      const long long rt = tile / n_blocks;
      const int nb = static_cast<int>(tile - rt * n_blocks);
      const long long p0 = rt * kMlpBM;
      const int n0 = nb * a.bn;
      const int bn = min(a.bn, a.n_pad - n0);
      const unsigned buf = nbuf == 2 ? static_cast<unsigned>(j & 1) : 0u;
      mbar_wait(&ctl.acc_full[buf], static_cast<unsigned>((nbuf == 2 ? (j >> 1) : j) & 1));
      tc_fence_after();
      const long long prow = p0 + warp * 32 + lane;
      const uint32_t stg = ring + static_cast<uint32_t>(S) * stage_bytes + warp * 4096u;  // after the ring
      const uint32_t lane_addr = tmem + buf * static_cast<uint32_t>(a.tmem_cols) + ((warp * 32u) << 16);
      // per-frame bias (a 128-row tile never straddles batch elements: bias_npb % 128 == 0)
      const float *bias_t = a.bias + (a.bias_npb > 0 ? (p0 / a.bias_npb) * a.n_pad : 0);
      for (int c0 = 0; c0 < bn; c0 += 32) {
        float v[32];
        const int cw = min(32, bn - c0);
        const unsigned chunk = lane & 7u;                       // STORE: the lane's group of four columns
        const bool col_on = static_cast<int>(chunk) * 4 < cw;
        float4 bq = make_float4(0.f, 0.f, 0.f, 0.f);
        if (EPI == EPI_STORE && col_on) bq = ldg128(bias_t + n0 + c0 + chunk * 4);
        float bch = 0.f;   // MAXPOOL_T: the lane's channel = n0 + 128 * (c0 / 128) + 32 * warp + lane
        if (EPI == EPI_MAXPOOL_T || EPI == EPI_STORE_T) bch = __ldg(bias_t + n0 + (c0 & ~127) + warp * 32 + lane);
        if (cw == 32) tmem_ld32(lane_addr + c0, v);
        else tmem_ld16(lane_addr + c0, v);
        if (EPI == EPI_SUMPOOL) {
          // sum over the 32 rows of the warp of relu(acc + bias): partial sums of a mean over points
          // (DenseFusion's AvgPool1d, pvn3d.py:165,178); rows past the end contribute 0
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bq = q * 4 < cw ? ldg128(bias_t + n0 + c0 + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            const bool on = prow < a.rows && q * 4 < cw;
            v[q * 4 + 0] = on ? fmaxf(v[q * 4 + 0] + bq.x, 0.f) : 0.f;
            v[q * 4 + 1] = on ? fmaxf(v[q * 4 + 1] + bq.y, 0.f) : 0.f;
            v[q * 4 + 2] = on ? fmaxf(v[q * 4 + 2] + bq.z, 0.f) : 0.f;
            v[q * 4 + 3] = on ? fmaxf(v[q * 4 + 3] + bq.w, 0.f) : 0.f;
          }
          warp_colsum_32(v, lane);
          const int col = static_cast<int>(lane);
          if (col < cw && p0 + warp * 32 < a.rows)
            a.out[((p0 + warp * 32) / 32) * a.ldo + a.col0 + n0 + c0 + col] = v[0];
        } else if (EPI == EPI_STORE_T) {
          // TRANSPOSED accumulator, stored channel-major: lane = channel, registers = 32 consecutive points of one
          // frame = 128 contiguous bytes of out[frame][channel][:] -- the [B, C, N] layout Pointnet2MSG.forward returns
          // (pvn3d.py:154) without a transposing pass over the [B*N, C] rows.  Bias / ReLU on the lane's channel, then
          // through the swizzled staging tile (row = channel) so that one STG.128 writes four full 128-byte lines
          // (direct stores -- 32 half sectors per instruction -- cost the gathering producers 64 us of LSU time)
          const long long prow0 = p0 + (c0 & 127);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 r = make_float4(v[q * 4 + 0] + bch, v[q * 4 + 1] + bch, v[q * 4 + 2] + bch, v[q * 4 + 3] + bch);
            if (a.relu) {
              r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
            }
            if (a.round_out) {
              r.x = to_tf32(r.x); r.y = to_tf32(r.y); r.z = to_tf32(r.z); r.w = to_tf32(r.w);
            }
            sts128(stg + lane * 128u + ((static_cast<uint32_t>(q) ^ (lane & 7u)) << 4), r.x, r.y, r.z, r.w);
          }
          __syncwarp();
          if (prow0 < a.rows) {
            const long long fb = prow0 / a.out_cn;
            const long long pt = prow0 - fb * a.out_cn;
            const int ch0 = n0 + (c0 & ~127) + static_cast<int>(warp * 32 + (lane >> 3));
            float *o = a.out + (fb * a.n_pad + ch0) * a.out_cn + pt + chunk * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const unsigned row = 4u * i + (lane >> 3);
              const float4 r = lds128(stg + row * 128u + ((chunk ^ (row & 7u)) << 4));
              *reinterpret_cast<float4 *>(o + static_cast<size_t>(4 * i) * a.out_cn) = r;
            }
          }
          __syncwarp();
        } else if (EPI == EPI_MAXPOOL_T) {
          // TRANSPOSED accumulator (the MMA ran as W . A^T): TMEM lane = output channel, column = row of the tile, so
          // the max over the `pool` consecutive rows of a centre is a max over REGISTERS of one thread -- no
          // shuffles -- and the 32 lanes of a warp store 32 consecutive channels of one pooled row (128 bytes).
          // Groups never straddle the end: rows % pool == 0 and 128 % pool == 0.
          const long long prow0 = p0 + (c0 & 127);          // row of register 0
          float *o = a.out + a.col0 + n0 + (c0 & ~127) + warp * 32 + lane;
          if (a.pool == 32) {
#pragma unroll
            for (int w = 16; w >= 1; w >>= 1)
#pragma unroll
              for (int i = 0; i < w; ++i) v[i] = fmaxf(v[i], v[i + w]);
            if (prow0 < a.rows) {
              float r = v[0] + bch;
              if (a.relu) r = fmaxf(r, 0.f);
              if (a.round_out) r = to_tf32(r);
              o[(prow0 >> 5) * a.ldo] = r;
            }
          } else if (a.pool == 16) {
#pragma unroll
            for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
              for (int i = 0; i < w; ++i) {
                v[i] = fmaxf(v[i], v[i + w]);
                v[16 + i] = fmaxf(v[16 + i], v[16 + i + w]);
              }
#pragma unroll
            for (int g = 0; g < 2; ++g)
              if (prow0 + 16 * g < a.rows) {
                float r = v[16 * g] + bch;
                if (a.relu) r = fmaxf(r, 0.f);
                if (a.round_out) r = to_tf32(r);
                o[((prow0 >> 4) + g) * a.ldo] = r;
              }
          } else {  // pool == 8
#pragma unroll
            for (int w = 4; w >= 1; w >>= 1)
#pragma unroll
              for (int i = 0; i < w; ++i)
#pragma unroll
                for (int g = 0; g < 4; ++g) v[8 * g + i] = fmaxf(v[8 * g + i], v[8 * g + i + w]);
#pragma unroll
            for (int g = 0; g < 4; ++g)
              if (prow0 + 8 * g < a.rows) {
                float r = v[8 * g] + bch;
                if (a.relu) r = fmaxf(r, 0.f);
                if (a.round_out) r = to_tf32(r);
                o[((prow0 >> 3) + g) * a.ldo] = r;
              }
          }
        } else if (EPI == EPI_STORE) {
          // raw accumulators through a swizzled 4 KB staging tile (thread = row going in, 8 lanes per row coming
          // out: the global stores are 128-byte row segments, 4 rows per STG.128); bias / ReLU / rounding on the
          // way out, where a lane keeps ONE group of four columns -- its bias is a single float4 (loaded before
          // the accumulator wait) and the eight rows it stores are independent instruction streams
          if (cw == 32) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              sts128(stg + lane * 128u + ((static_cast<uint32_t>(q) ^ (lane & 7u)) << 4), v[q * 4 + 0], v[q * 4 + 1],
                     v[q * 4 + 2], v[q * 4 + 3]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              sts128(stg + lane * 128u + ((static_cast<uint32_t>(q) ^ (lane & 7u)) << 4), v[q * 4 + 0], v[q * 4 + 1],
                     v[q * 4 + 2], v[q * 4 + 3]);
          }
          __syncwarp();
          if (col_on) {
            const bool full = p0 + kMlpBM <= a.rows;
            float4 r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const unsigned row = 4u * i + (lane >> 3);
              r[i] = lds128(stg + row * 128u + ((chunk ^ (row & 7u)) << 4));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              r[i].x += bq.x; r[i].y += bq.y; r[i].z += bq.z; r[i].w += bq.w;
              if (a.relu) {
                r[i].x = fmaxf(r[i].x, 0.f); r[i].y = fmaxf(r[i].y, 0.f); r[i].z = fmaxf(r[i].z, 0.f); r[i].w = fmaxf(r[i].w, 0.f);
              }
              if (a.round_out) {
                r[i].x = to_tf32(r[i].x); r[i].y = to_tf32(r[i].y); r[i].z = to_tf32(r[i].z); r[i].w = to_tf32(r[i].w);
              }
            }
            float *o = a.out + (p0 + warp * 32 + (lane >> 3)) * a.ldo + a.col0 + n0 + c0 + chunk * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (full || p0 + warp * 32 + 4 * i + (lane >> 3) < a.rows)
                *reinterpret_cast<float4 *>(o + static_cast<size_t>(4 * i) * a.ldo) = r[i];
          }
          __syncwarp();
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
              float r = v[0] + __ldg(bias_t + n0 + c0 + col);
              if (a.relu) r = fmaxf(r, 0.f);
              if (a.round_out) r = to_tf32(r);
              a.out[grow0 * a.ldo + a.col0 + n0 + c0 + col] = r;
            }
          } else if (a.pool == 16) {
            warp_colmax_16(v, lane);
            const int g = lane >> 4, col = static_cast<int>(lane & 15u) * 2;
            if (col < cw && p0 + warp * 32 + g * 16 < a.rows) {
              float2 r;
              r.x = v[0] + __ldg(bias_t + n0 + c0 + col);
              r.y = v[1] + __ldg(bias_t + n0 + c0 + col + 1);
              if (a.relu) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); }
              if (a.round_out) { r.x = to_tf32(r.x); r.y = to_tf32(r.y); }
              *reinterpret_cast<float2 *>(a.out + (grow0 + g) * a.ldo + a.col0 + n0 + c0 + col) = r;
            }
          } else {  // pool == 8
            warp_colmax_8(v, lane);
            const int g = lane >> 3, col = static_cast<int>(lane & 7u) * 4;
            if (col < cw && p0 + warp * 32 + g * 8 < a.rows) {
              float4 r;
              r.x = v[0] + __ldg(bias_t + n0 + c0 + col);
              r.y = v[1] + __ldg(bias_t + n0 + c0 + col + 1);
              r.z = v[2] + __ldg(bias_t + n0 + c0 + col + 2);
              r.w = v[3] + __ldg(bias_t + n0 + c0 + col + 3);
              if (a.relu) {
                r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
              }
              if (a.round_out) {
                r.x = to_tf32(r.x); r.y = to_tf32(r.y); r.z = to_tf32(r.z); r.w = to_tf32(r.w);
              }
              *reinterpret_cast<float4 *>(a.out + (grow0 + g) * a.ldo + a.col0 + n0 + c0 + col) = r;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&ctl.acc_empty[buf]);  // accumulator `buf` may be overwritten
    }
  } else {
    // ================= warp 12: MMA issuer =============================================================
    long long it_base = 0, j = 0;
    for (long long tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, it_base += kc_total, ++j) {
      const long long rt = tile / n_blocks;
      const int nb = static_cast<int>(tile - rt * n_blocks);
      const int bn = min(a.bn, a.n_pad - nb * a.bn);
      const uint32_t idesc = instr_desc_tf32(bn);
      const unsigned buf = nbuf == 2 ? static_cast<unsigned>(j & 1) : 0u;
      mbar_wait(&ctl.acc_empty[buf], static_cast<unsigned>(((nbuf == 2 ? (j >> 1) : j) & 1) ^ 1));
      tc_fence_after();
      const uint32_t acc = tmem + buf * static_cast<uint32_t>(a.tmem_cols);
      for (int kc = 0; kc < kc_total; ++kc) {
        const long long it = it_base + kc;
        const int s = static_cast<int>(it % S);
        mbar_wait(&ctl.full[s], static_cast<unsigned>((it / S) & 1));
        fence_proxy_async_smem();  // cp.async (generic proxy) data of the stage -> async proxy
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = ring + static_cast<uint32_t>(s) * stage_bytes;
          const uint64_t adesc = smem_desc_sw128(sa), bdesc = smem_desc_sw128(sa + a_bytes);
          if (EPI == EPI_MAXPOOL_T || EPI == EPI_STORE_T) {
            // operands swapped: D^T[channel][row] = W . A^T, one M = 128 block of channels per accumulator of 128
            // columns (the tile's rows); W's next 128 rows are 16 KB (>> 4 = 1024) further
            const uint32_t idesc_t = instr_desc_tf32(kMlpBM);
            for (int h = 0; h < bn / 128; ++h)
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4)
                umma_tf32(acc + static_cast<uint32_t>(h * 128), bdesc + static_cast<uint64_t>(h * 1024 + k4 * 2),
                          adesc + static_cast<uint64_t>(k4 * 2), idesc_t, (kc > 0 || k4 > 0) ? 1u : 0u);
          } else {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)  // K = 8 tf32 = 32 bytes per instruction: +2 in 16-byte units
              umma_tf32(acc, adesc + static_cast<uint64_t>(k4 * 2), bdesc + static_cast<uint64_t>(k4 * 2),
                        idesc, (kc > 0 || k4 > 0) ? 1u : 0u);
          }
          umma_commit(&ctl.empty[s]);  // stage reusable once these MMAs have read it
          if (kc == kc_total - 1) umma_commit(&ctl.acc_full[buf]);
        }
        __syncwarp();
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == kMlpEpiWarps + kMlpProWarps) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                 "r"(nbuf * static_cast<uint32_t>(a.tmem_cols))
                 : "memory");
  }
}

bool weight_tensor_map(CUtensorMap *map, const float *w, int k_pad, int n_pad, int bn);   // below

template <int PRO, int EPI>
int launch_mlp(MlpArgs &a, cudaStream_t st) {
  if (a.rows <= 0) return PVN3D_OK;
  if (a.k_pad <= 0 || a.k_pad % 32 || a.n_pad <= 0 || a.n_pad % 16) return PVN3D_ERR_INVALID_ARG;
  // columns per tile: <= 256, multiple of 16, as even a split as possible
  const int nblk = ceil_div(a.n_pad, 256);
  a.bn = ((ceil_div(a.n_pad, nblk) + 15) / 16) * 16;
  int tc = 32;
  while (tc < a.bn) tc <<= 1;
  a.tmem_cols = tc;
  const size_t stage_bytes = kMlpBM * 128 + align_up(static_cast<size_t>(a.bn) * 128, 1024);
  const int sms = std::max(1, sm_count() - a.reserve_sms);
  const long long tiles = ((a.rows + kMlpBM - 1) / kMlpBM) * ceil_div(a.n_pad, a.bn);
  // two CTAs per SM when both fit: accumulators <= 512 TMEM columns in total (2 x 2 x <=128, or 2 x 1 x 256: wide
  // tiles give up the second accumulator of the CTA, the co-resident CTA fills the gap), >= min_stages stages in
  // half the shared memory, and enough tiles to feed twice the CTAs (PVN3D_MLP_OCC=1 forces one CTA per SM,
  // PVN3D_MLP_ACC1=0 keeps wide tiles at one CTA per SM, PVN3D_MLP_OCC2_TILES = tiles-per-SM threshold)
  static const int occ_env = [] { const char *e = getenv("PVN3D_MLP_OCC"); return e ? atoi(e) : 0; }();
  static const int acc1_env = [] { const char *e = getenv("PVN3D_MLP_ACC1"); return e ? atoi(e) : 1; }();
  static const int tiles_env = [] { const char *e = getenv("PVN3D_MLP_OCC2_TILES"); return e ? atoi(e) : 4; }();
  static const int budget_env = [] { const char *e = getenv("PVN3D_MLP_OCC2_KB"); return e ? atoi(e) : 96; }();
  const size_t budget2 = static_cast<size_t>(std::min(std::max(budget_env, 48), 96)) * 1024;   // ring of one of two co-resident CTAs
  // asynchronous producers (pre-rounded dense activations) keep two chunks in flight: >= 3 stages
  const size_t min_stages = ((PRO == PRO_DENSE || PRO == PRO_SA_GATHER) && a.a_tf32) ? 3 : 2;
  // ... and at two CTAs per SM a ring of only three stages starves them unless a tile is a single chunk (measured:
  // 96 -> 128 max-pool layer 154 us with 6 stages at one CTA per SM, 213 us with 3 stages at two)
  const size_t min_stages2 = (min_stages == 3 && a.k_pad > 32) ? 4 : min_stages;
  const bool fits2 = min_stages2 * stage_bytes <= budget2 && tiles >= static_cast<long long>(tiles_env) * sms;
  bool occ2 = a.tmem_cols <= 128 && fits2;
  a.acc_bufs = 2;
  if (!occ2 && a.tmem_cols == 256 && fits2 && acc1_env) {
    occ2 = true;
    a.acc_bufs = 1;
  }
  if (occ_env == 1) {
    occ2 = false;
    a.acc_bufs = 2;
  }
  int stages = static_cast<int>((occ2 ? budget2 : size_t(208 * 1024)) / stage_bytes);
  if (stages > kMlpMaxStages) stages = kMlpMaxStages;
  if (stages < 2) stages = 2;
  a.stages = stages;
  const size_t smem = stages * stage_bytes + kMlpEpiWarps * 4096 + 256;  // ring + epilogue staging + barriers
  static_assert(sizeof(MlpSmemCtl) <= 256, "barrier block");
  a.use_tma = weight_tensor_map(&a.tmap, a.w, a.k_pad, a.n_pad, a.bn) ? 1 : 0;
  if (occ2) {
    auto kern = mlp_layer_kernel<PRO, EPI, 2>;
    static PerDeviceOnce once2;
    PVN3D_ONCE_PER_DEVICE(once2,
                          (cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100),
                           cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024)),
                          "mlp smem attr (2 CTAs/SM)");
    const unsigned grid = static_cast<unsigned>(std::min<long long>(tiles, 2ll * sms));
    kern<<<grid, kMlpThreads, smem, st>>>(a);
    return check_launch("mlp_layer_kernel<2>");
  }
  auto kern = mlp_layer_kernel<PRO, EPI, 1>;
  static PerDeviceOnce once;
  PVN3D_ONCE_PER_DEVICE(once,
                        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024),
                        "mlp smem attr");
  const unsigned grid = static_cast<unsigned>(std::min<long long>(tiles, sms));
  kern<<<grid, kMlpThreads, smem, st>>>(a);
  return check_launch("mlp_layer_kernel");
}

// =====================================================================================================
// The whole SharedMLP of one SA scale (3 layers) / FP module (2 layers) in ONE persistent kernel.
//
// One launch per layer sends every inter-layer activation through HBM twice (write + read: 7.0 of the
// 10.4 GB a 32-frame batch moves, 6.8 GB measured by ncu in round 1).  Rows of a shared MLP are
// independent, so a CTA can take a 128-row tile through ALL layers: layer l's epilogue stores its
// TF32-rounded tile into a CTA-private scratch tile (128 x n_pad floats, re-used for every row tile of the
// CTA: it lives in L2 and is overwritten before it is ever evicted), and the producers of layer l+1 pull
// it back with cp.async.  To keep the tensor core, the producers and the epilogue busy while one tile
// waits on its own previous layer, a CTA works on TWO row tiles at a time ("slots"), in the order
//     (A,L0) (B,L0) (A,L1) (B,L1) (A,L2) (B,L2) | next pair ...
// Every (tile, layer, column block) is one ITEM; the three roles (producers / MMA issuer / epilogue) walk
// the same item sequence, exactly like the per-layer kernel walks tiles: the operand ring and the two
// TMEM accumulators are shared by all layers.  New dependency: the producers of (slot, l>0) wait on
// h_ready[slot], on which the 128 epilogue threads arrive after storing (slot, l-1).
constexpr int kChainMaxSlots = 8;
struct ChainLayer {
  const float *w, *bias;
  int k_pad, n_pad, bn, n_blocks;
};
struct MlpChainArgs {
  // weight matrices as TMA tensor maps: [n_pad][k_pad] fp32, box = 32 columns (128 B) x bn rows,
  // SWIZZLE_128B -- one cp.async.bulk.tensor.2d per K chunk lands the B operand in the UMMA layout
  alignas(64) CUtensorMap tmap[3];
  int use_tma;
  MlpArgs base;        // producer of layer 0 + final epilogue (rows, out, ldo, col0, pool)
  ChainLayer layer[3];
  int n_layers;
  float *scratch;      // [grid][n_slots][slot_floats]
  int h_off[2];        // float offset of layer l's output tile inside a slot (l < n_layers - 1)
  int slot_floats;
  int n_slots;         // row tiles a CTA keeps in flight (2..kChainMaxSlots)
};

struct ChainSmemCtl {
  uint64_t full[kMlpMaxStages];
  uint64_t empty[kMlpMaxStages];
  uint64_t acc_full[2];
  uint64_t acc_empty[2];
  uint64_t h_ready[kChainMaxSlots];   // 128 arrivals: epilogue threads, after storing a slot's intermediate tile
  uint64_t h_seen[kChainMaxSlots];    // 256 arrivals: every producer thread, once it has passed a h_ready phase
  uint32_t tmem_base;
};

template <int PRO, int EPI>
__global__ void __launch_bounds__(kMlpThreads, 1) mlp_chain_kernel(const __grid_constant__ MlpChainArgs c) {
  extern __shared__ unsigned char mlp_smem_raw[];
  __shared__ ChainSmemCtl ctl;
  const MlpArgs &a = c.base;
  const uint32_t raw = smem_u32(mlp_smem_raw);
  const uint32_t ring = (raw + 1023u) & ~1023u;
  const uint32_t a_bytes = kMlpBM * 128u;
  int bn_max = 0;
  for (int l = 0; l < c.n_layers; ++l) bn_max = max(bn_max, c.layer[l].bn);
  const uint32_t stage_bytes = a_bytes + ((static_cast<uint32_t>(bn_max) * 128u + 1023u) & ~1023u);

  const int t = threadIdx.x;
  const unsigned warp = t >> 5, lane = t & 31u;
  const int S = a.stages;
  const int NL = c.n_layers;
  const long long row_tiles = (a.rows + kMlpBM - 1) / kMlpBM;
  // row tiles of this CTA: blockIdx.x, +grid, ...; handled NS at a time
  const int NS = c.n_slots;
  const long long my_tiles = row_tiles > blockIdx.x ? (row_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long n_pairs = (my_tiles + NS - 1) / NS;
  float *const scratch = c.scratch + static_cast<size_t>(blockIdx.x) * NS * c.slot_floats;

  if (warp == kMlpEpiWarps + kMlpProWarps) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(&ctl.full[s], 128);
        mbar_init(&ctl.empty[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&ctl.acc_full[b], 1);
        mbar_init(&ctl.acc_empty[b], 128);
      }
      for (int b = 0; b < kChainMaxSlots; ++b) {
        mbar_init(&ctl.h_ready[b], 128);
        mbar_init(&ctl.h_seen[b], kMlpProWarps * 32);
      }
      mbar_fence_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&ctl.tmem_base)),
                 "r"(static_cast<uint32_t>(2 * a.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ctl.tmem_base;

  if (warp >= kMlpEpiWarps && warp < kMlpEpiWarps + kMlpProWarps) {
    // ================= producers ======================================================================
    const int pt = (t - kMlpEpiWarps * 32) & 127;
    const unsigned grp = (warp - kMlpEpiWarps) >> 2;
    const int pw = pt >> 5, sub = static_cast<int>(lane & 7u), rg = static_cast<int>(lane >> 3);
    const int r_first = 32 * pw + rg;
    bool vec_ok = true;
    if (PRO == PRO_DENSE) vec_ok = (a.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.a) & 15u) == 0);
    if (PRO == PRO_SA_GATHER)
      vec_ok = a.c_feat > 0 && (a.ldf % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.feat) & 15u) == 0);
    if (PRO == PRO_FP_INTERP)
      vec_ok = (a.c2 % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.known_feat) & 15u) == 0);
    int pend0 = 0, pend1 = 0, npend = 0;
    unsigned hwaits = 0u;   // bit `slot`: parity of the next h_ready phase this thread waits for
    long long it_base = 0;
    for (long long pair = 0; pair < n_pairs; ++pair) {
      for (int l = 0; l < NL; ++l) {
        const ChainLayer &L = c.layer[l];
        const int kc_total = L.k_pad / 32;
        for (int slot = 0; slot < NS; ++slot) {
          const long long local = NS * pair + slot;
          if (local >= my_tiles) continue;
          const long long rt = blockIdx.x + local * gridDim.x;
          const long long p_first = rt * kMlpBM + r_first;
          RowState rs;
          int a_cols_l = 0;
          if (l == 0) {
            rows_setup<PRO>(a, p_first, rs);
          } else {
            // input = this slot's intermediate tile of layer l-1 (all 128 rows exist, TF32-rounded).
            // Nothing of this thread may stay unpublished while it blocks: the MMA warp consumes chunks in
            // order, and the tile awaited here is produced behind every chunk staged so far.
            if (npend > 0) {
              cp_async_wait<0>();
              fence_proxy_async_smem();
              mbar_arrive(&ctl.full[pend0]);
              if (npend == 2) mbar_arrive(&ctl.full[pend1]);
              npend = 0;
            }
            mbar_wait(&ctl.h_ready[slot], (hwaits >> slot) & 1u);
            hwaits ^= 1u << slot;
            mbar_arrive(&ctl.h_seen[slot]);   // this phase has been observed: the epilogue may open the next one
            const float *h = scratch + static_cast<size_t>(slot) * c.slot_floats + c.h_off[l - 1];
            const int ld = c.layer[l - 1].n_pad;
            rs.live = 0xffu;
#pragma unroll
            for (int j = 0; j < 8; ++j) rs.row[j] = h + static_cast<size_t>(r_first + 4 * j) * ld;
            a_cols_l = ld;
          }
          for (int nb = 0; nb < L.n_blocks; ++nb) {
            const int n0 = nb * L.bn;
            const int bn = min(L.bn, L.n_pad - n0);
            for (int kc = 0; kc < kc_total; ++kc) {
              const long long it = it_base + kc;
              if (static_cast<unsigned>(it & 1) != grp) continue;
              const int s = static_cast<int>(it % S);
              if (npend == 2 || (l == 0 && npend > 0)) {
                // publish the older asynchronous chunk(s) before this thread can block on a free stage
                // (layer 0 stages synchronously: everything still pending is flushed first)
                if (npend == 2 && l != 0) {
                  cp_async_wait<1>();
                  fence_proxy_async_smem();
                  mbar_arrive(&ctl.full[pend0]);
                  pend0 = pend1;
                  npend = 1;
                } else {
                  cp_async_wait<0>();
                  fence_proxy_async_smem();
                  mbar_arrive(&ctl.full[pend0]);
                  if (npend == 2) mbar_arrive(&ctl.full[pend1]);
                  npend = 0;
                }
              }
              mbar_wait(&ctl.empty[s], static_cast<unsigned>(((it / S) & 1) ^ 1));
              const uint32_t sa = ring + static_cast<uint32_t>(s) * stage_bytes;
              const uint32_t sb = sa + a_bytes;
              if (c.use_tma) {
                // B operand by the TMA engine: rows n0..n0+L.bn of W, columns kc*32..+32, swizzled on the fly;
                // completes on the stage's `full` barrier (transaction bytes registered first)
                if (pt == 0) {
                  mbar_expect_tx_only(&ctl.full[s], static_cast<unsigned>(L.bn) * 128u);
                  tma_load_2d(sb, &c.tmap[l], kc * 32, n0, &ctl.full[s]);
                }
              } else {
                for (int i = pt; i < bn * 8; i += 128) {
                  const int n = i >> 3, cc = i & 7;
                  cp_async16(sb + sw128_off(n, cc), L.w + static_cast<size_t>(n0 + n) * L.k_pad + kc * 32 + cc * 4);
                }
              }
              if (l > 0) {
                const int k = kc * 32 + 4 * sub;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  cp_async16(sa + sw128_off(r_first + 4 * j, sub), rs.row[j] + (k < a_cols_l ? k : 0),
                             k < a_cols_l ? 16u : 0u);
                cp_async_commit();
                if (npend == 0) pend0 = s; else pend1 = s;
                ++npend;
              } else {
                cp_async_commit();
                stage_a_chunk<PRO>(a, rs, p_first, r_first, sub, kc * 32, sa, vec_ok);
                cp_async_wait<0>();
                fence_proxy_async_smem();
                mbar_arrive(&ctl.full[s]);
              }
            }
            it_base += kc_total;
          }
        }
      }
    }
    if (npend > 0) {
      cp_async_wait<0>();
      fence_proxy_async_smem();
      mbar_arrive(&ctl.full[pend0]);
      if (npend == 2) mbar_arrive(&ctl.full[pend1]);
    }
  } else if (warp < kMlpEpiWarps) {
    // ================= epilogue =========================================================================
    long long j = 0;
    unsigned hsig_any = 0u, hsig_par = 0u;   // per slot: signalled before / parity of the last phase signalled
    const uint32_t stg = ring + static_cast<uint32_t>(S) * stage_bytes + warp * 4096u;
    for (long long pair = 0; pair < n_pairs; ++pair) {
      for (int l = 0; l < NL; ++l) {
        const ChainLayer &L = c.layer[l];
        const bool last = l == NL - 1;
        for (int slot = 0; slot < NS; ++slot) {
          const long long local = NS * pair + slot;
          if (local >= my_tiles) continue;
          const long long rt = blockIdx.x + local * gridDim.x;
          const long long p0 = rt * kMlpBM;
          float *const h = scratch + static_cast<size_t>(slot) * c.slot_floats + (last ? 0 : c.h_off[l]);
          for (int nb = 0; nb < L.n_blocks; ++nb, ++j) {
            const int n0 = nb * L.bn;
            const int bn = min(L.bn, L.n_pad - n0);
            const unsigned buf = static_cast<unsigned>(j & 1);
            mbar_wait(&ctl.acc_full[buf], static_cast<unsigned>((j >> 1) & 1));
            tc_fence_after();
            const long long prow = p0 + warp * 32 + lane;
            const uint32_t lane_addr = tmem + buf * static_cast<uint32_t>(a.tmem_cols) + ((warp * 32u) << 16);
            for (int c0 = 0; c0 < bn; c0 += 32) {
              float v[32];
              const int cw = min(32, bn - c0);
              if (cw == 32) tmem_ld32(lane_addr + c0, v);
              else tmem_ld16(lane_addr + c0, v);
              if (!last || EPI == EPI_STORE) {
                // intermediate: bias + ReLU + TF32 rounding into the slot's scratch tile (row = local row);
                // final STORE: bias + ReLU (+ rounding if asked) into out
                const bool rnd = !last || a.round_out;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  if (q * 4 < cw) {
                    const float4 bq = ldg128(L.bias + n0 + c0 + q * 4);
                    float4 r;
                    r.x = fmaxf(v[q * 4 + 0] + bq.x, 0.f);
                    r.y = fmaxf(v[q * 4 + 1] + bq.y, 0.f);
                    r.z = fmaxf(v[q * 4 + 2] + bq.z, 0.f);
                    r.w = fmaxf(v[q * 4 + 3] + bq.w, 0.f);
                    if (rnd) {
                      r.x = to_tf32(r.x); r.y = to_tf32(r.y); r.z = to_tf32(r.z); r.w = to_tf32(r.w);
                    }
                    sts128(stg + lane * 128u + ((static_cast<uint32_t>(q) ^ (lane & 7u)) << 4), r.x, r.y, r.z, r.w);
                  }
                }
                __syncwarp();
                const unsigned chunk = lane & 7u;
                if (static_cast<int>(chunk) * 4 < cw) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    const unsigned row = 4u * i + (lane >> 3);
                    const float4 r = lds128(stg + row * 128u + ((chunk ^ (row & 7u)) << 4));
                    if (!last) {
                      *reinterpret_cast<float4 *>(h + static_cast<size_t>(warp * 32 + row) * L.n_pad + n0 + c0 + chunk * 4) = r;
                    } else {
                      const long long pr = p0 + warp * 32 + row;
                      if (pr < a.rows)
                        *reinterpret_cast<float4 *>(a.out + pr * a.ldo + a.col0 + n0 + c0 + chunk * 4) = r;
                    }
                  }
                }
                __syncwarp();
              } else {
                // final layer of an SA scale: ReLU(max over the `pool` rows of each centre + bias)
                if (prow >= a.rows) {
#pragma unroll
                  for (int i = 0; i < 32; ++i) v[i] = -__int_as_float(0x7f800000);
                }
                const long long grow0 = (p0 + warp * 32) / a.pool;
                if (a.pool == 32) {
                  warp_colmax_32(v, lane);
                  const int col = static_cast<int>(lane);
                  if (col < cw && p0 + warp * 32 < a.rows)
                    a.out[grow0 * a.ldo + a.col0 + n0 + c0 + col] = fmaxf(v[0] + __ldg(L.bias + n0 + c0 + col), 0.f);
                } else if (a.pool == 16) {
                  warp_colmax_16(v, lane);
                  const int g = lane >> 4, col = static_cast<int>(lane & 15u) * 2;
                  if (col < cw && p0 + warp * 32 + g * 16 < a.rows) {
                    float2 r;
                    r.x = fmaxf(v[0] + __ldg(L.bias + n0 + c0 + col), 0.f);
                    r.y = fmaxf(v[1] + __ldg(L.bias + n0 + c0 + col + 1), 0.f);
                    *reinterpret_cast<float2 *>(a.out + (grow0 + g) * a.ldo + a.col0 + n0 + c0 + col) = r;
                  }
                } else {  // pool == 8
                  warp_colmax_8(v, lane);
                  const int g = lane >> 3, col = static_cast<int>(lane & 7u) * 4;
                  if (col < cw && p0 + warp * 32 + g * 8 < a.rows) {
                    float4 r;
                    r.x = fmaxf(v[0] + __ldg(L.bias + n0 + c0 + col), 0.f);
                    r.y = fmaxf(v[1] + __ldg(L.bias + n0 + c0 + col + 1), 0.f);
                    r.z = fmaxf(v[2] + __ldg(L.bias + n0 + c0 + col + 2), 0.f);
                    r.w = fmaxf(v[3] + __ldg(L.bias + n0 + c0 + col + 3), 0.f);
                    *reinterpret_cast<float4 *>(a.out + (grow0 + g) * a.ldo + a.col0 + n0 + c0 + col) = r;
                  }
                }
              }
            }
            tc_fence_before();
            mbar_arrive(&ctl.acc_empty[buf]);
          }
          if (!last) {
            // the slot's tile of layer l is complete in global memory (L2): make it visible at GPU scope
            // (the producers read it back with cp.async.cg, which goes to L2) and release the next layer
            __threadfence();
            // phase k of h_ready[slot] may only complete once every producer thread has observed phase k-1
            // (a parity wait cannot tell phase k-1 from phase k+1)
            if ((hsig_any >> slot) & 1u) mbar_wait(&ctl.h_seen[slot], (hsig_par >> slot) & 1u);
            else hsig_par |= 1u << slot;       // first signal: the next wait is for phase 0 -> parity 0 after the toggle
            hsig_any |= 1u << slot;
            hsig_par ^= 1u << slot;
            mbar_arrive(&ctl.h_ready[slot]);
          }
        }
      }
    }
  } else {
    // ================= warp 12: MMA issuer =============================================================
    long long it_base = 0, j = 0;
    for (long long pair = 0; pair < n_pairs; ++pair) {
      for (int l = 0; l < NL; ++l) {
        const ChainLayer &L = c.layer[l];
        const int kc_total = L.k_pad / 32;
        for (int slot = 0; slot < NS; ++slot) {
          if (NS * pair + slot >= my_tiles) continue;
          for (int nb = 0; nb < L.n_blocks; ++nb, ++j, it_base += kc_total) {
            const int bn = min(L.bn, L.n_pad - nb * L.bn);
            const uint32_t idesc = instr_desc_tf32(bn);
            const unsigned buf = static_cast<unsigned>(j & 1);
            mbar_wait(&ctl.acc_empty[buf], static_cast<unsigned>(((j >> 1) & 1) ^ 1));
            tc_fence_after();
            const uint32_t acc = tmem + buf * static_cast<uint32_t>(a.tmem_cols);
            for (int kc = 0; kc < kc_total; ++kc) {
              const long long it = it_base + kc;
              const int s = static_cast<int>(it % S);
              mbar_wait(&ctl.full[s], static_cast<unsigned>((it / S) & 1));
              fence_proxy_async_smem();
              tc_fence_after();
              if (lane == 0) {
                const uint32_t sa = ring + static_cast<uint32_t>(s) * stage_bytes;
                const uint64_t adesc = smem_desc_sw128(sa), bdesc = smem_desc_sw128(sa + a_bytes);
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                  umma_tf32(acc, adesc + static_cast<uint64_t>(k4 * 2), bdesc + static_cast<uint64_t>(k4 * 2),
                            idesc, (kc > 0 || k4 > 0) ? 1u : 0u);
                umma_commit(&ctl.empty[s]);
                if (kc == kc_total - 1) umma_commit(&ctl.acc_full[buf]);
              }
              __syncwarp();
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == kMlpEpiWarps + kMlpProWarps) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem),
                 "r"(static_cast<uint32_t>(2 * a.tmem_cols))
                 : "memory");
  }
}

// per-layer tile geometry of a chain; returns the scratch floats one slot needs
int chain_plan(MlpChainArgs &c, const pvn3d_mlp_layer_t *layers, int n_layers) {
  int bn_max = 0, off = 0;
  for (int l = 0; l < n_layers; ++l) {
    ChainLayer &L = c.layer[l];
    L.w = layers[l].w; L.bias = layers[l].bias; L.k_pad = layers[l].k_pad; L.n_pad = layers[l].n_pad;
    L.n_blocks = ceil_div(L.n_pad, 256);
    L.bn = ((ceil_div(L.n_pad, L.n_blocks) + 15) / 16) * 16;
    bn_max = std::max(bn_max, L.bn);
    if (l < n_layers - 1) {
      c.h_off[l] = off;
      off += kMlpBM * L.n_pad;
    }
  }
  c.n_layers = n_layers;
  c.slot_floats = off;
  int tc = 32;
  while (tc < bn_max) tc <<= 1;
  c.base.tmem_cols = tc;
  c.base.bn = bn_max;
  return off;
}

// cuTensorMapEncodeTiled lives in libcuda; the library links only the runtime, so the entry point is
// fetched once through cudaGetDriverEntryPoint
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    tried = true;
  }
  return fn;
}

// tensor map of one weight matrix; false -> the kernel falls back to per-thread cp.async
bool weight_tensor_map(CUtensorMap *map, const float *w, int k_pad, int n_pad, int bn) {
  const char *env = getenv("PVN3D_MLP_TMA");
  if (env && env[0] == '0') return false;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc || (reinterpret_cast<uintptr_t>(w) & 15u)) return false;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(k_pad), static_cast<cuuint64_t>(n_pad)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(k_pad) * sizeof(float)};
  const cuuint32_t box[2] = {32u, static_cast<cuuint32_t>(bn)};
  const cuuint32_t estr[2] = {1u, 1u};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(w), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
bool chain_tensor_maps(MlpChainArgs &c) {
  for (int l = 0; l < c.n_layers; ++l)
    if (!weight_tensor_map(&c.tmap[l], c.layer[l].w, c.layer[l].k_pad, c.layer[l].n_pad, c.layer[l].bn)) return false;
  return true;
}

bool chain_layers_ok(const pvn3d_mlp_layer_t *layers, int n_layers, int k_first_min) {
  if (!layers || n_layers < 1 || n_layers > 3) return false;
  for (int l = 0; l < n_layers; ++l) {
    if (!layers[l].w || !layers[l].bias || layers[l].k_pad <= 0 || layers[l].k_pad % 32 ||
        layers[l].n_pad <= 0 || layers[l].n_pad % 16)
      return false;
    // layer l reads the n_pad(l-1) columns of the previous tile: its K must cover them
    if (l > 0 && layers[l].k_pad < layers[l - 1].n_pad) return false;
  }
  return layers[0].k_pad >= k_first_min;
}

// row tiles a CTA keeps in flight: enough to cover the epilogue -> L2 -> producer latency of a layer
// transition, as long as all scratch tiles of the grid stay well inside the 126 MB L2 (<= 48 MB)
int chain_slots(int slot_floats, long long row_tiles, unsigned grid) {
  int want = 4;
  if (const char *env = getenv("PVN3D_CHAIN_SLOTS")) want = atoi(env);
  want = std::max(2, std::min(kChainMaxSlots, want));
  const long long per_cta = (row_tiles + grid - 1) / std::max(1u, grid);
  while (want > 2 && want > per_cta) --want;
  while (want > 2 && static_cast<size_t>(grid) * want * slot_floats * sizeof(float) > (48u << 20)) --want;
  return want;
}

template <int PRO, int EPI>
int launch_chain(MlpChainArgs &c, void *workspace, size_t workspace_bytes, cudaStream_t st) {
  MlpArgs &a = c.base;
  if (a.rows <= 0) return PVN3D_OK;
  const size_t stage_bytes = kMlpBM * 128 + align_up(static_cast<size_t>(a.bn) * 128, 1024);
  int stages = static_cast<int>((208 * 1024) / stage_bytes);
  if (stages > kMlpMaxStages) stages = kMlpMaxStages;
  if (stages < 3) return PVN3D_ERR_UNSUPPORTED;   // asynchronous producers keep two chunks in flight
  a.stages = stages;
  const size_t smem = stages * stage_bytes + 1024 + kMlpEpiWarps * 4096;
  auto kern = mlp_chain_kernel<PRO, EPI>;
  static PerDeviceOnce once;
  PVN3D_ONCE_PER_DEVICE(once,
                        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024),
                        "mlp chain smem attr");
  const int sms = std::max(1, sm_count() - a.reserve_sms);
  const long long row_tiles = (a.rows + kMlpBM - 1) / kMlpBM;
  const unsigned grid = static_cast<unsigned>(std::min<long long>(row_tiles, sms));
  c.n_slots = chain_slots(c.slot_floats, row_tiles, grid);
  const size_t need = static_cast<size_t>(grid) * c.n_slots * c.slot_floats * sizeof(float);
  if (c.n_layers > 1 && (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 15u)))
    return PVN3D_ERR_WORKSPACE;
  c.scratch = static_cast<float *>(workspace);
  c.use_tma = chain_tensor_maps(c) ? 1 : 0;
  kern<<<grid, kMlpThreads, smem, st>>>(c);
  return check_launch("mlp_chain_kernel");
}

int dispatch(MlpArgs &a, int pro, int pool, cudaStream_t st) {
  if (pool) {
    if (pool != 8 && pool != 16 && pool != 32) return PVN3D_ERR_UNSUPPORTED;
    a.pool = pool;
    // 128-channel pooled layers: transposed accumulator, register max instead of shuffles (79 -> 51 us and 154 -> 125 us
    // on the SA2 scales).  Tiles of 256 channels work too (PVN3D_MLP_POOLT=2) but measured 5-25 % SLOWER: two
    // M=128 x N=128 TF32 MMAs per K step read 128 B of shared memory per clock, the N=256 form 96.  PVN3D_MLP_POOLT=0:
    // the shuffle epilogue everywhere.
    static const int poolt_env = [] { const char *e = getenv("PVN3D_MLP_POOLT"); return e ? atoi(e) : 1; }();
    const bool t128 = a.n_pad == 128;
    const bool t256 = a.n_pad % 128 == 0 && (a.n_pad <= 256 || a.n_pad % 256 == 0);
    if (pro == PRO_DENSE && ((poolt_env == 1 && t128) || (poolt_env == 2 && t256)))
      return launch_mlp<PRO_DENSE, EPI_MAXPOOL_T>(a, st);
    if (pro == PRO_DENSE) return launch_mlp<PRO_DENSE, EPI_MAXPOOL>(a, st);
    if (pro == PRO_SA_GATHER) return launch_mlp<PRO_SA_GATHER, EPI_MAXPOOL>(a, st);
    if (pro == PRO_SA_FACT) return launch_mlp<PRO_SA_FACT, EPI_MAXPOOL>(a, st);
    return PVN3D_ERR_INVALID_ARG;
  }
  a.pool = 0;
  if (pro == PRO_DENSE) return launch_mlp<PRO_DENSE, EPI_STORE>(a, st);
  if (pro == PRO_SA_FACT) return launch_mlp<PRO_SA_FACT, EPI_STORE>(a, st);
  if (pro == PRO_FP_FACT) return launch_mlp<PRO_FP_FACT, EPI_STORE>(a, st);
  if (pro == PRO_SA_GATHER) return launch_mlp<PRO_SA_GATHER, EPI_STORE>(a, st);
  return launch_mlp<PRO_FP_INTERP, EPI_STORE>(a, st);
}

// ---- factored first layer of an SA scale -------------------------------------------------------------
// table row j = [ tf32(f_j) (C) | hi(x_j) (3) | lo(x_j) (3) | 0.. ]: x = hi + lo with hi = tf32(x), lo = tf32(x - hi),
// so that a TF32 GEMM against [W_f | W_x | W_x] evaluates W_x . x to ~2^-21 relative (the coordinates are
// ~1 m and their DIFFERENCES ~1 cm: a single TF32 rounding of x would cost 10 % of the difference)
__global__ void sa_factor_table_kernel(const float *__restrict__ xyz, const float *__restrict__ feat, int ldf,
                                       int c_feat, long long rows, int k_pad, float *__restrict__ out) {
  // one 16-byte group of a row per thread (k_pad / 4 threads per row): 128-byte stores, 4+ rows per warp
  const int qpr = k_pad >> 2;
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long p = g / qpr;
  if (p >= rows) return;
  const int c0 = static_cast<int>(g - p * qpr) * 4;
  float v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = c0 + e;
    float r = 0.f;
    if (c < c_feat) {
      r = to_tf32(__ldg(feat + p * ldf + c));
    } else if (c < c_feat + 6) {
      const int d = (c - c_feat) % 3;
      const float x = __ldg(xyz + p * 3 + d);
      const float hi = to_tf32(x);
      r = c < c_feat + 3 ? hi : to_tf32(x - hi);
    }
    v[e] = r;
  }
  *reinterpret_cast<float4 *>(out + p * k_pad + c0) = make_float4(v[0], v[1], v[2], v[3]);
}
// V[i, n] = sum_d Wx[n, d] * c_i[d] - bias[n]   (fp32 FMAs; Wx = the TF32-rounded xyz columns of W1)
__global__ void sa_centre_term_kernel(const float *__restrict__ centres, const float *__restrict__ wx,
                                      const float *__restrict__ bias, long long rows, int n_pad,
                                      float *__restrict__ out) {
  const long long p = static_cast<long long>(blockIdx.x) * blockDim.y + threadIdx.y;
  if (p >= rows) return;
  const float cx = __ldg(centres + p * 3), cy = __ldg(centres + p * 3 + 1), cz = __ldg(centres + p * 3 + 2);
  for (int n = threadIdx.x; n < n_pad; n += blockDim.x) {
    const float w0 = __ldg(wx + n * 3), w1 = __ldg(wx + n * 3 + 1), w2 = __ldg(wx + n * 3 + 2);
    out[p * n_pad + n] = __fmaf_rn(w2, cz, __fmaf_rn(w1, cy, w0 * cx)) - __ldg(bias + n);
  }
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
                               const float *bias, int k_pad, int n_pad, int flags, int pool,
                               float *out, int ldo, int col0, pvn3d_stream_t stream) {
  if (!a || !w || !bias || !out || lda < a_cols || a_cols < 0 || a_cols % 4 || rows < 0 || ldo % 4 ||
      col0 % 4)
    return PVN3D_ERR_INVALID_ARG;
  if (pool && rows % pool) return PVN3D_ERR_INVALID_ARG;
  MlpArgs m{};
  m.w = w; m.bias = bias; m.rows = rows; m.k_pad = k_pad; m.n_pad = n_pad;
  m.a = a; m.lda = lda; m.a_cols = a_cols;
  m.out = out; m.ldo = ldo; m.col0 = col0;
  m.relu = flags & PVN3D_MLP_RELU; m.round_out = (flags & PVN3D_MLP_ROUND_OUT) ? 1 : 0;
  m.a_tf32 = (flags & PVN3D_MLP_A_TF32) ? 1 : 0;
  m.reserve_sms = (flags >> 8) & 0xff;
  return dispatch(m, PRO_DENSE, pool, as_stream(stream));
}

extern "C" int pvn3d_mlp_sa_first(const float *xyz, const float *new_xyz, const float *feat_pm,
                                  int ldf, int c_feat, const int *idx, int b, int n, int m, int ns,
                                  const float *w, const float *bias, int k_pad, int n_pad, int flags,
                                  int pool, float *out, int ldo, int col0, pvn3d_stream_t stream) {
  if (!xyz || !new_xyz || !idx || !w || !bias || !out || b < 0 || n <= 0 || m < 0 || ns <= 0 ||
      c_feat < 0 || (c_feat > 0 && (!feat_pm || ldf < c_feat)) || k_pad < c_feat + 3 || ldo % 4 ||
      col0 % 4)
    return PVN3D_ERR_INVALID_ARG;
  if (pool && pool != ns) return PVN3D_ERR_INVALID_ARG;
  if (static_cast<long long>(m) * ns > 0x3fffffffll || static_cast<long long>(b) * n > 0x7fffffffll ||
      static_cast<long long>(b) * m > 0x7fffffffll)
    return PVN3D_ERR_UNSUPPORTED;
  MlpArgs a{};
  a.w = w; a.bias = bias; a.rows = static_cast<long long>(b) * m * ns; a.k_pad = k_pad; a.n_pad = n_pad;
  a.xyz = xyz; a.new_xyz = new_xyz; a.feat = feat_pm; a.ldf = ldf; a.c_feat = c_feat; a.idx = idx;
  a.n = n; a.m = m; a.ns = ns;
  a.out = out; a.ldo = ldo; a.col0 = col0;
  a.relu = flags & PVN3D_MLP_RELU; a.round_out = (flags & PVN3D_MLP_ROUND_OUT) ? 1 : 0;
  a.a_tf32 = (flags & PVN3D_MLP_A_TF32) ? 1 : 0;   // feat_pm rows are TF32-rounded: asynchronous copies
  a.reserve_sms = (flags >> 8) & 0xff;
  return dispatch(a, PRO_SA_GATHER, pool, as_stream(stream));
}

extern "C" int pvn3d_mlp_fp_first(const float *known_feat_pm, int c2, const int *nn_idx,
                                  const float *nn_w, const float *skip_pm, int lds, int c1, int b,
                                  int n_unknown, int m_known, const float *w, const float *bias,
                                  int k_pad, int n_pad, int flags, float *out, int ldo, int col0,
                                  pvn3d_stream_t stream) {
  if (!known_feat_pm || !nn_idx || !nn_w || !w || !bias || !out || b < 0 || n_unknown < 0 ||
      m_known <= 0 || c2 <= 0 || c1 < 0 || (c1 > 0 && (!skip_pm || lds < c1)) || k_pad < c2 + c1 ||
      ldo % 4 || col0 % 4)
    return PVN3D_ERR_INVALID_ARG;
  if (static_cast<long long>(b) * m_known > 0x7fffffffll || n_unknown > 0x3fffffff)
    return PVN3D_ERR_UNSUPPORTED;
  MlpArgs a{};
  a.w = w; a.bias = bias; a.rows = static_cast<long long>(b) * n_unknown; a.k_pad = k_pad; a.n_pad = n_pad;
  a.known_feat = known_feat_pm; a.c2 = c2; a.nn_idx = nn_idx; a.nn_w = nn_w; a.skip = skip_pm;
  a.lds = lds; a.c1 = c1; a.n_unknown = n_unknown; a.m_known = m_known;
  a.out = out; a.ldo = ldo; a.col0 = col0;
  a.relu = flags & PVN3D_MLP_RELU; a.round_out = (flags & PVN3D_MLP_ROUND_OUT) ? 1 : 0;
  a.reserve_sms = (flags >> 8) & 0xff;
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

extern "C" size_t pvn3d_mlp_chain_workspace_bytes(const pvn3d_mlp_layer_t *layers, int n_layers) {
  if (!layers || n_layers < 1 || n_layers > 3) return 0;
  size_t floats = 0;
  for (int l = 0; l + 1 < n_layers; ++l) floats += static_cast<size_t>(kMlpBM) * layers[l].n_pad;
  return static_cast<size_t>(std::max(1, sm_count())) * kChainMaxSlots * floats * sizeof(float) + 256;
}

extern "C" int pvn3d_mlp_sa_chain(const float *xyz, const float *new_xyz, const float *feat_pm, int ldf,
                                  int c_feat, const int *idx, int b, int n, int m, int ns,
                                  const pvn3d_mlp_layer_t *layers, int n_layers, int flags, int pool,
                                  float *out, int ldo, int col0, void *workspace, size_t workspace_bytes,
                                  pvn3d_stream_t stream) {
  if (!xyz || !new_xyz || !idx || !out || b < 0 || n <= 0 || m < 0 || ns <= 0 || c_feat < 0 ||
      (c_feat > 0 && (!feat_pm || ldf < c_feat)) || ldo % 4 || col0 % 4 ||
      !chain_layers_ok(layers, n_layers, c_feat + 3))
    return PVN3D_ERR_INVALID_ARG;
  if (pool && pool != ns) return PVN3D_ERR_INVALID_ARG;
  if (pool && pool != 8 && pool != 16 && pool != 32) return PVN3D_ERR_UNSUPPORTED;
  if (static_cast<long long>(m) * ns > 0x3fffffffll || static_cast<long long>(b) * n > 0x7fffffffll ||
      static_cast<long long>(b) * m > 0x7fffffffll)
    return PVN3D_ERR_UNSUPPORTED;
  MlpChainArgs c{};
  MlpArgs &a = c.base;
  a.rows = static_cast<long long>(b) * m * ns;
  a.xyz = xyz; a.new_xyz = new_xyz; a.feat = feat_pm; a.ldf = ldf; a.c_feat = c_feat; a.idx = idx;
  a.n = n; a.m = m; a.ns = ns;
  a.out = out; a.ldo = ldo; a.col0 = col0; a.relu = 1; a.round_out = 0; a.pool = pool;
  a.reserve_sms = (flags >> 8) & 0xff;
  a.k_pad = layers[0].k_pad;
  chain_plan(c, layers, n_layers);
  if (pool) return launch_chain<PRO_SA_GATHER, EPI_MAXPOOL>(c, workspace, workspace_bytes, as_stream(stream));
  return launch_chain<PRO_SA_GATHER, EPI_STORE>(c, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int pvn3d_mlp_fp_chain(const float *known_feat_pm, int c2, const int *nn_idx, const float *nn_w,
                                  const float *skip_pm, int lds, int c1, int b, int n_unknown,
                                  int m_known, const pvn3d_mlp_layer_t *layers, int n_layers, int flags,
                                  float *out, int ldo, int col0, void *workspace, size_t workspace_bytes,
                                  pvn3d_stream_t stream) {
  if (!known_feat_pm || !nn_idx || !nn_w || !out || b < 0 || n_unknown < 0 || m_known <= 0 || c2 <= 0 ||
      c1 < 0 || (c1 > 0 && (!skip_pm || lds < c1)) || ldo % 4 || col0 % 4 ||
      !chain_layers_ok(layers, n_layers, c2 + c1))
    return PVN3D_ERR_INVALID_ARG;
  if (static_cast<long long>(b) * m_known > 0x7fffffffll || n_unknown > 0x3fffffff)
    return PVN3D_ERR_UNSUPPORTED;
  MlpChainArgs c{};
  MlpArgs &a = c.base;
  a.rows = static_cast<long long>(b) * n_unknown;
  a.known_feat = known_feat_pm; a.c2 = c2; a.nn_idx = nn_idx; a.nn_w = nn_w; a.skip = skip_pm;
  a.lds = lds; a.c1 = c1; a.n_unknown = n_unknown; a.m_known = m_known;
  a.out = out; a.ldo = ldo; a.col0 = col0; a.relu = 1; a.round_out = 0; a.pool = 0;
  a.reserve_sms = (flags >> 8) & 0xff;
  a.k_pad = layers[0].k_pad;
  chain_plan(c, layers, n_layers);
  return launch_chain<PRO_FP_INTERP, EPI_STORE>(c, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int pvn3d_sa_factor_table(const float *xyz, const float *feat_pm, int ldf, int c_feat, long long rows,
                                     int k_pad, float *out, pvn3d_stream_t stream) {
  if (!xyz || !out || rows < 0 || c_feat < 0 || (c_feat > 0 && (!feat_pm || ldf < c_feat)) || k_pad < c_feat + 6 ||
      k_pad % 32)
    return PVN3D_ERR_INVALID_ARG;
  if (rows == 0) return PVN3D_OK;
  const long long groups = rows * (k_pad / 4);
  if ((groups + 255) / 256 > 0x7fffffffll || (reinterpret_cast<uintptr_t>(out) & 15u)) return PVN3D_ERR_UNSUPPORTED;
  sa_factor_table_kernel<<<static_cast<unsigned>((groups + 255) / 256), 256, 0, as_stream(stream)>>>(xyz, feat_pm, ldf, c_feat,
                                                                                                    rows, k_pad, out);
  return check_launch("sa_factor_table_kernel");
}

extern "C" int pvn3d_sa_centre_term(const float *centres, const float *wx, const float *bias, long long rows,
                                    int n_pad, float *out, pvn3d_stream_t stream) {
  if (!centres || !wx || !bias || !out || rows < 0 || n_pad <= 0 || n_pad % 4) return PVN3D_ERR_INVALID_ARG;
  if (rows == 0) return PVN3D_OK;
  const dim3 block(32, 8);
  sa_centre_term_kernel<<<static_cast<unsigned>((rows + 7) / 8), block, 0, as_stream(stream)>>>(centres, wx, bias, rows,
                                                                                               n_pad, out);
  return check_launch("sa_centre_term_kernel");
}

extern "C" int pvn3d_mlp_sa_fact(const float *u, const float *v, int ldu, int c_valid, const int *idx, int b, int n,
                                 int m, int ns, const float *w, const float *bias, int k_pad, int n_pad, int flags,
                                 int pool, float *out, int ldo, int col0, pvn3d_stream_t stream) {
  if (!u || !v || !idx || !w || !bias || !out || b < 0 || n <= 0 || m < 0 || ns <= 0 || c_valid <= 0 ||
      c_valid % 4 || ldu < c_valid || ldu % 4 || k_pad < c_valid || ldo % 4 || col0 % 4 ||
      (reinterpret_cast<uintptr_t>(u) & 15u) || (reinterpret_cast<uintptr_t>(v) & 15u))
    return PVN3D_ERR_INVALID_ARG;
  if (pool && pool != ns) return PVN3D_ERR_INVALID_ARG;
  if (static_cast<long long>(m) * ns > 0x3fffffffll || static_cast<long long>(b) * n > 0x7fffffffll ||
      static_cast<long long>(b) * m > 0x7fffffffll)
    return PVN3D_ERR_UNSUPPORTED;
  MlpArgs a{};
  a.w = w; a.bias = bias; a.rows = static_cast<long long>(b) * m * ns; a.k_pad = k_pad; a.n_pad = n_pad;
  a.feat = u; a.new_xyz = v; a.ldf = ldu; a.c_feat = c_valid; a.idx = idx;
  a.n = n; a.m = m; a.ns = ns;
  a.out = out; a.ldo = ldo; a.col0 = col0;
  a.relu = flags & PVN3D_MLP_RELU; a.round_out = (flags & PVN3D_MLP_ROUND_OUT) ? 1 : 0;
  a.reserve_sms = (flags >> 8) & 0xff;
  return dispatch(a, PRO_SA_FACT, pool, as_stream(stream));
}

extern "C" int pvn3d_mlp_fp_fact(const float *p, const float *s, int ld, int c_valid, const int *nn_idx,
                                 const float *nn_w, int b, int n_unknown, int m_known, const float *w,
                                 const float *bias, int k_pad, int n_pad, int flags, float *out, int ldo, int col0,
                                 pvn3d_stream_t stream) {
  if (!p || !s || !nn_idx || !nn_w || !w || !bias || !out || b < 0 || n_unknown < 0 || m_known <= 0 || c_valid <= 0 ||
      c_valid % 4 || ld != c_valid || k_pad < c_valid || ldo % 4 || col0 % 4 || (reinterpret_cast<uintptr_t>(p) & 15u) ||
      (reinterpret_cast<uintptr_t>(s) & 15u))
    return PVN3D_ERR_INVALID_ARG;
  if (static_cast<long long>(b) * m_known > 0x7fffffffll || n_unknown > 0x3fffffff) return PVN3D_ERR_UNSUPPORTED;
  MlpArgs a{};
  a.w = w; a.bias = bias; a.rows = static_cast<long long>(b) * n_unknown; a.k_pad = k_pad; a.n_pad = n_pad;
  a.known_feat = p; a.c2 = ld; a.nn_idx = nn_idx; a.nn_w = nn_w; a.skip = s; a.lds = ld; a.c1 = 0;
  a.n_unknown = n_unknown; a.m_known = m_known;
  a.out = out; a.ldo = ldo; a.col0 = col0;
  a.relu = flags & PVN3D_MLP_RELU; a.round_out = (flags & PVN3D_MLP_ROUND_OUT) ? 1 : 0;
  a.reserve_sms = (flags >> 8) & 0xff;
  if (flags & PVN3D_MLP_OUT_CN) {
    // out = [b][n_pad][n_unknown]: one or two 128-channel accumulators per tile, 32-point groups inside a frame
    if ((n_pad != 128 && n_pad != 256) || n_unknown % 32 || (reinterpret_cast<uintptr_t>(out) & 15u))
      return PVN3D_ERR_UNSUPPORTED;
    a.out_cn = n_unknown; a.pool = 0; a.ldo = n_pad; a.col0 = 0;
    return launch_mlp<PRO_FP_FACT, EPI_STORE_T>(a, as_stream(stream));
  }
  return dispatch(a, PRO_FP_FACT, 0, as_stream(stream));
}

extern "C" int pvn3d_mlp_dense_frame_bias(const float *a, int lda, int a_cols, long long rows, int rows_per_frame,
                                          const float *w, const float *bias, int k_pad, int n_pad, int flags,
                                          float *out, int ldo, int col0, pvn3d_stream_t stream) {
  if (!a || !w || !bias || !out || lda < a_cols || a_cols < 0 || a_cols % 4 || rows < 0 || ldo % 4 || col0 % 4 ||
      rows_per_frame <= 0 || rows_per_frame % kMlpBM || rows % rows_per_frame)
    return PVN3D_ERR_INVALID_ARG;
  MlpArgs m{};
  m.w = w; m.bias = bias; m.rows = rows; m.k_pad = k_pad; m.n_pad = n_pad;
  m.a = a; m.lda = lda; m.a_cols = a_cols;
  m.out = out; m.ldo = ldo; m.col0 = col0;
  m.relu = flags & PVN3D_MLP_RELU; m.round_out = (flags & PVN3D_MLP_ROUND_OUT) ? 1 : 0;
  m.a_tf32 = (flags & PVN3D_MLP_A_TF32) ? 1 : 0;
  m.reserve_sms = (flags >> 8) & 0xff;
  m.bias_npb = rows_per_frame;
  return dispatch(m, PRO_DENSE, 0, as_stream(stream));
}

extern "C" int pvn3d_mlp_dense_sum32(const float *a, int lda, int a_cols, long long rows, const float *w,
                                     const float *bias, int k_pad, int n_pad, int flags, float *out, int ldo,
                                     int col0, pvn3d_stream_t stream) {
  if (!a || !w || !bias || !out || lda < a_cols || a_cols < 0 || a_cols % 4 || rows < 0 || ldo % 4 || col0 % 4)
    return PVN3D_ERR_INVALID_ARG;
  MlpArgs m{};
  m.w = w; m.bias = bias; m.rows = rows; m.k_pad = k_pad; m.n_pad = n_pad;
  m.a = a; m.lda = lda; m.a_cols = a_cols;
  m.out = out; m.ldo = ldo; m.col0 = col0;
  m.relu = 1;
  m.a_tf32 = (flags & PVN3D_MLP_A_TF32) ? 1 : 0;
  m.reserve_sms = (flags >> 8) & 0xff;
  m.pool = 32;
  return launch_mlp<PRO_DENSE, EPI_SUMPOOL>(m, as_stream(stream));
}
