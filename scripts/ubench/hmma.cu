// Micro-benchmark: legacy warp-level mma.sync (m16n8k16, fp16 in, fp32 accumulate) issue cost per SM sub-partition on
// B200, alone and interleaved with MUFU / FFMA2 work (does it run on a pipe of its own?).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  uint32_t a[4] = {0x3c003c00u + threadIdx.x, 0x3c003800u, 0x38003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x38003800u + threadIdx.x};
  float d[8][4];
  float m[8];
  u64 p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { m[i] = 1.f + i + threadIdx.x * 1e-3f; p[i] = (u64)__float_as_uint(m[i]) << 32 | __float_as_uint(m[i]);
    for (int j = 0; j < 4; ++j) d[i][j] = 0.f; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      hmma(d[i], a, b);
      if (MODE >= 1) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(m[i])); asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(m[i])); }
      if (MODE >= 2) { asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[i])); asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[(i + 3) & 7]));
                       asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[(i + 5) & 7])); }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3] + m[i] + __uint_as_float((uint32_t)p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char* name, int w) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 1000;
  k<MODE><<<148, w * 128>>>(out, cyc, 10);
  k<MODE><<<148, w * 128>>>(out, cyc, iters);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s warps/sched=%d  cycles per group per scheduler = %.2f\n", name, w, (double)h / ((double)iters * 8 * w));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int w : {1, 2, 4}) {
    if (w == 1) { run<0>("HMMA.16816.F32", 1); run<1>("HMMA + EX2 + SQRT (XU floor 16)", 1); run<2>("HMMA + EX2 + SQRT + 3 FFMA2 (floors 16 / 6)", 1); }
    if (w == 2) { run<0>("HMMA.16816.F32", 2); run<1>("HMMA + EX2 + SQRT (XU floor 16)", 2); run<2>("HMMA + EX2 + SQRT + 3 FFMA2 (floors 16 / 6)", 2); }
    if (w == 4) { run<0>("HMMA.16816.F32", 4); run<1>("HMMA + EX2 + SQRT (XU floor 16)", 4); run<2>("HMMA + EX2 + SQRT + 3 FFMA2 (floors 16 / 6)", 4); }
  }
  return 0;
}
