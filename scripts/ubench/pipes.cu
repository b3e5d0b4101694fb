// Micro-benchmark: issue cost (cycles per warp instruction per SM sub-partition) of the instructions the scoring
// kernel's residual warps are made of.  One CTA per SM, W warps per scheduler, long unrolled independent chains.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }

template <int OP>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 1.0f + threadIdx.x * 1e-3f + i;
  u64 p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = pk2(a[i], a[i] + 0.5f);
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 2) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 3) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[i]));
      if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (OP == 5) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 6) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 7) { asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[i]));
                     asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[(i + 1) & 7])); asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[(i + 2) & 7])); }
      if (OP == 8) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[i]));
                     asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[(i + 1) & 7])); asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p[(i + 2) & 7])); }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { float lo, hi; upk2(p[i], lo, hi); s += a[i] + lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, int warps_per_sched, int per_iter) {
  float* out; long long* cyc;
  const int threads = warps_per_sched * 4 * 32, iters = 2000;
  cudaMalloc(&out, 148 * threads * 4); cudaMalloc(&cyc, 8);
  k<OP><<<148, threads>>>(out, cyc, 10);
  k<OP><<<148, threads>>>(out, cyc, iters);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double inst_per_sched = (double)iters * 8 * per_iter * warps_per_sched;
  printf("%-28s warps/sched=%d  cycles per warp-instruction per scheduler = %.2f\n", name, warps_per_sched, h / inst_per_sched);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {1, 2, 4, 8}) {
    if (w == 1) { run<0>("MUFU.SQRT", 1, 1); run<1>("MUFU.EX2", 1, 1); run<2>("MUFU.RSQ", 1, 1); run<3>("FFMA2", 1, 1); run<4>("FFMA", 1, 1); run<5>("MUFU.RCP", 1, 1); run<6>("MUFU.LG2", 1, 1); run<7>("SQRT+3xFFMA2 (per 4 inst)", 1, 4); run<8>("EX2+3xFFMA2 (per 4 inst)", 1, 4); }
    if (w == 2) { run<0>("MUFU.SQRT", 2, 1); run<1>("MUFU.EX2", 2, 1); run<2>("MUFU.RSQ", 2, 1); run<3>("FFMA2", 2, 1); run<4>("FFMA", 2, 1); run<7>("SQRT+3xFFMA2 (per 4 inst)", 2, 4); run<8>("EX2+3xFFMA2 (per 4 inst)", 2, 4); }
    if (w == 4) { run<0>("MUFU.SQRT", 4, 1); run<1>("MUFU.EX2", 4, 1); run<2>("MUFU.RSQ", 4, 1); run<3>("FFMA2", 4, 1); run<4>("FFMA", 4, 1); run<7>("SQRT+3xFFMA2 (per 4 inst)", 4, 4); run<8>("EX2+3xFFMA2 (per 4 inst)", 4, 4); }
    if (w == 8) { run<0>("MUFU.SQRT", 8, 1); run<1>("MUFU.EX2", 8, 1); run<2>("MUFU.RSQ", 8, 1); run<3>("FFMA2", 8, 1); run<4>("FFMA", 8, 1); run<7>("SQRT+3xFFMA2 (per 4 inst)", 8, 4); run<8>("EX2+3xFFMA2 (per 4 inst)", 8, 4); }
  }
  return 0;
}
