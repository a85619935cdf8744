// micro-benchmark: how MUFU.EX2 overlaps with FP32-pipe work (FFMA / FFMA2) on one SM sub-partition
#include <cstdio>
#include <cuda_runtime.h>
template <int NF, int PACKED, int NM>
__global__ void __launch_bounds__(256) k(float *out, int iters) {
  float2 acc[8];
  float e[4];
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i);
  for (int i = 0; i < 4; ++i) e[i] = threadIdx.x * 1e-3f + i;
  const float2 m = make_float2(0.999f, 1.001f), c = make_float2(1e-3f, 2e-3f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NF; ++i) {
      if (PACKED) acc[i & 7] = __ffma2_rn(acc[i & 7], m, c);
      else acc[i & 7].x = __fmaf_rn(acc[i & 7].x, m.x, c.x);
    }
#pragma unroll
    for (int i = 0; i < NM; ++i) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(e[i & 3])); e[i & 3] = y; }
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  for (int i = 0; i < 4; ++i) s += e[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NF, int PACKED, int NM>
void run(const char *name) {
  float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  const int iters = 20000;
  k<NF, PACKED, NM><<<148 * 8, 256>>>(out, 10);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<NF, PACKED, NM><<<148 * 8, 256>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double thr_iters = 148.0 * 8 * 256 * (double)iters;
  const double clk = ms * 1e-3 * 1.9e9 * 148;  // SM-cycles at 1.9 GHz
  printf("%-34s %.3f ms | FP lane-ops %.1f /clk/SM | MUFU %.1f /clk/SM | SMSP-cycles per warp-iteration %.1f\n", name, ms,
         thr_iters * NF * (PACKED ? 2 : 1) / clk, thr_iters * NM / clk, clk * 4 / (thr_iters / 32));
}
int main() {
  run<8, 0, 0>("8 FFMA");
  run<8, 1, 0>("8 FFMA2");
  run<0, 0, 1>("1 MUFU");
  run<0, 0, 2>("2 MUFU");
  run<8, 0, 1>("8 FFMA + 1 MUFU");
  run<4, 1, 1>("4 FFMA2 + 1 MUFU (old sweep)");
  run<2, 1, 1>("2 FFMA2 + 1 MUFU (mma sweep)");
  run<4, 0, 1>("4 FFMA + 1 MUFU");
  run<16, 0, 1>("16 FFMA + 1 MUFU");
  return 0;
}
