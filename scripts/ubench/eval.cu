// Micro-benchmark of the residual tail (residual_tail of score_tc.cu: everything after u = R n^ for one hypothesis row x one
// column pair: 13 packed FMA-pipe instructions + 4 MUFU.SQRT + 4 MUFU.EX2 + 2 cvt) in isolation: cycles per evaluation per SM
// sub-partition for 1/2/4/8 warps per scheduler, constants in registers.  (profiles/ubench_eval_all_fma.txt is the same
// measurement for the earlier all-FMA evaluation, 25 packed instructions: 66-67 cycles at >= 2 warps per scheduler.)
#include <cstdio>
#include "../../nopesac_b200/csrc/score_tc.cu"

__global__ void k_eval(const float* in, uint32_t* out, long long* cyc, int iters) {
  float R[9], tr[3];
  for (int i = 0; i < 9; ++i) R[i] = in[i] + threadIdx.x * 1e-4f;
  for (int i = 0; i < 3; ++i) tr[i] = in[9 + i];
  u64 c[CJ8_FIELDS];
  for (int i = 0; i < CJ8_FIELDS; ++i) c[i] = pk2(in[12 + i] + (threadIdx.x & 31) * 1e-3f, in[24 + i]);
  uint32_t acc = 0;
  u64 s0 = 0, s1 = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t hr, ht;
      R[0] += 1e-6f;
      residual_tail<false>(pk2(R[0], R[1]), pk2(R[2], R[3]), pk2(R[4], R[5]), pk2(tr[0], tr[1]), c, 0ull, hr, ht, s0, s1);
      acc ^= hr + ht;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float h[40];
  for (int i = 0; i < 40; ++i) h[i] = 0.1f * (i % 7) + 0.05f;
  float* in; uint32_t* out; long long* cyc;
  cudaMalloc(&in, sizeof(h)); cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int w : {1, 2, 4, 8}) {
    const int iters = 500;
    k_eval<<<148, w * 128>>>(in, out, cyc, 10);
    k_eval<<<148, w * 128>>>(in, out, cyc, iters);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("residual_tail: warps/sched=%d  cycles per evaluation per scheduler = %.1f  (XU floor 64, FMA-pipe floor 26)\n", w,
           (double)c / ((double)iters * 8 * w));
  }
  return 0;
}
