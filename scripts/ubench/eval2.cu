// Micro-benchmark, second stage: the residual k-block loop of score_tc.cu (all-FMA path, lane = column pair) in
// isolation, adding one ingredient of the real kernel at a time:  MODE 0 = math only (constants in registers),
// 1 = + row constants from shared memory (3 LDS.128 per evaluation), 2 = + the two STS.32 per evaluation,
// 3 = + fence.proxy.async + __syncwarp + mbarrier arrive per 8 evaluations, 4 = + an mbarrier try_wait per k-block.
#include <cstdio>
#include "../../nopesac_b200/csrc/score_tc.cu"

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_eval(const float* in, uint32_t* out, long long* cyc, int iters) {
  extern __shared__ uint8_t sm[];
  float* rowc = reinterpret_cast<float*>(sm) + (threadIdx.x >> 5) * 96;     // [warp][8][12]
  uint8_t* atile = sm + 16384;                                              // 32 KB A tile pair
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 16384 + 32768);
  const int lane = threadIdx.x & 31, rw = (threadIdx.x >> 5) & 15;
  if (threadIdx.x == 0) { mbar_init(bar, blockDim.x / 32); mbar_init(bar + 1, 1); }
  for (int i = lane; i < 96; i += 32) rowc[i] = in[i % 12] + i * 1e-3f;
  float R[9], tr[3];
  for (int i = 0; i < 9; ++i) R[i] = in[i] + threadIdx.x * 1e-4f;
  for (int i = 0; i < 3; ++i) tr[i] = in[9 + i];
  u64 c[CJ_FIELDS];
  for (int i = 0; i < CJ_FIELDS; ++i) c[i] = pk2(in[12 + i] + lane * 1e-3f, in[24 + i]);
  uint32_t acc = 0;
  u64 s0 = 0, s1 = 0;
  __syncthreads();
  uint32_t ph = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE >= 4) { uint32_t done; asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar + 1)), "r"(1u) : "memory"); acc += done; }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t hr, ht;
      if (MODE >= 1) {
        const float4 r0 = *reinterpret_cast<const float4*>(rowc + i * 12), r1 = *reinterpret_cast<const float4*>(rowc + i * 12 + 4),
                     r2 = *reinterpret_cast<const float4*>(rowc + i * 12 + 8);
        const float Rr[9] = {r0.x + it * 1e-6f, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x};
        const float tt[3] = {r2.y, r2.z, r2.w};
        residual_cp<false>(Rr, tt, c, hr, ht, s0, s1);
      } else {
        R[0] += 1e-6f;
        residual_cp<false>(R, tr, c, hr, ht, s0, s1);
      }
      if (MODE >= 2) {
        const int row = rw + 16 * i;
        const uint32_t off = (uint32_t)row * 128u + (uint32_t)(((lane >> 2) ^ (row & 7)) << 4) + (uint32_t)((lane & 3) << 2);
        *reinterpret_cast<uint32_t*>(atile + off) = hr;
        *reinterpret_cast<uint32_t*>(atile + 16384 + off) = ht;
      } else {
        acc ^= hr + ht;
      }
    }
    if (MODE >= 3) {
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const float* in, uint32_t* out, long long* cyc, int w) {
  const int iters = 500, smem = 16384 + 32768 + 64;
  cudaFuncSetAttribute(k_eval<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_eval<MODE><<<148, w * 128, smem>>>(in, out, cyc, 10);
  k_eval<MODE><<<148, w * 128, smem>>>(in, out, cyc, iters);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("mode %d warps/sched=%d  cycles per evaluation per scheduler = %.1f   (%s)\n", MODE, w, (double)c / ((double)iters * 8 * w),
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float h[40];
  for (int i = 0; i < 40; ++i) h[i] = 0.1f * (i % 7) + 0.05f;
  float* in; uint32_t* out; long long* cyc;
  cudaMalloc(&in, sizeof(h)); cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int w : {2, 4}) {
    run<0>(in, out, cyc, w); run<1>(in, out, cyc, w); run<2>(in, out, cyc, w); run<3>(in, out, cyc, w); run<4>(in, out, cyc, w);
  }
  return 0;
}
