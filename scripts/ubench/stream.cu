// Micro-benchmark: HBM streaming bandwidth reachable by ONE 256-thread "gather" role per SM (148 CTAs), as a function of
// the bytes in flight: (a) LDG.128 with U loads in flight per thread, (b) cp.async.bulk (TMA) into a shared-memory ring of
// S stages x 8 KB, consumed by the same 256 threads.  The scoring kernel's feature stream has exactly this shape.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int U>
__global__ void __launch_bounds__(256) k_ldg(const float4* __restrict__ src, float4* out, size_t n_per_cta) {
  const float4* p = src + (size_t)blockIdx.x * n_per_cta + threadIdx.x;
  float4 acc = make_float4(0, 0, 0, 0);
  for (size_t i = 0; i + (size_t)U * 256 <= n_per_cta; i += (size_t)U * 256) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(p + i + (size_t)u * 256);
#pragma unroll
    for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  }
  out[blockIdx.x * 256 + threadIdx.x] = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int S>
__global__ void __launch_bounds__(256) k_tma(const float4* __restrict__ src, float4* out, size_t n_per_cta) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S * 8192);
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int s = 0; s < 2 * S; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(s < S ? 1 : 256));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const char* base = reinterpret_cast<const char*>(src + (size_t)blockIdx.x * n_per_cta);
  const int nchunk = (int)(n_per_cta * 16 / 8192);
  float4 acc = make_float4(0, 0, 0, 0);
  if (tid == 0)
    for (int c = 0; c < S && c < nchunk; ++c) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[c])), "r"(8192) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + c * 8192)),
                   "l"(base + (size_t)c * 8192), "r"(8192), "r"(smem_u32(&bars[c])) : "memory");
    }
  for (int c = 0; c < nchunk; ++c) {
    const int s = c % S; const uint32_t ph = (c / S) & 1;
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bars[s])), "r"(ph) : "memory");
    const float4* v = reinterpret_cast<const float4*>(sm + s * 8192);
    const float4 a = v[tid], b = v[tid + 256];
    acc.x += a.x + b.x; acc.y += a.y + b.y; acc.z += a.z + b.z; acc.w += a.w + b.w;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[S + s])) : "memory");
    if (tid == 0 && c + S < nchunk) {
      done = 0;
      while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bars[S + s])), "r"(ph) : "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(8192) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + s * 8192)),
                   "l"(base + (size_t)(c + S) * 8192), "r"(8192), "r"(smem_u32(&bars[s])) : "memory");
    }
  }
  out[blockIdx.x * 256 + tid] = acc;
}

int main() {
  const size_t per_cta = (size_t)7 * 262144 / 16;      // 7 tiles x 262 KB per SM, in float4
  const size_t n = per_cta * 148;
  float4 *src, *out;
  cudaMalloc(&src, n * 16); cudaMalloc(&out, 148 * 256 * 16);
  cudaMemset(src, 0, n * 16);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time = [&](auto launch, const char* name) {
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    printf("%-34s %8.1f us  %7.0f GB/s  (%s)\n", name, ms * 1e3, n * 16 / (ms * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  };
  time([&] { k_ldg<4><<<148, 256>>>(src, out, per_cta); }, "LDG.128 x4 in flight (16 KB/SM)");
  time([&] { k_ldg<8><<<148, 256>>>(src, out, per_cta); }, "LDG.128 x8 in flight (32 KB/SM)");
  time([&] { k_ldg<12><<<148, 256>>>(src, out, per_cta); }, "LDG.128 x12 in flight (48 KB/SM)");
  time([&] { k_ldg<16><<<148, 256>>>(src, out, per_cta); }, "LDG.128 x16 in flight (64 KB/SM)");
  time([&] { k_ldg<24><<<148, 256>>>(src, out, per_cta); }, "LDG.128 x24 in flight (96 KB/SM)");
  cudaFuncSetAttribute(k_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
  cudaFuncSetAttribute(k_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
  cudaFuncSetAttribute(k_tma<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
  cudaFuncSetAttribute(k_tma<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
  time([&] { k_tma<2><<<148, 256, 2 * 8192 + 256>>>(src, out, per_cta); }, "TMA bulk ring 2 x 8 KB");
  time([&] { k_tma<4><<<148, 256, 4 * 8192 + 256>>>(src, out, per_cta); }, "TMA bulk ring 4 x 8 KB");
  time([&] { k_tma<8><<<148, 256, 8 * 8192 + 512>>>(src, out, per_cta); }, "TMA bulk ring 8 x 8 KB");
  time([&] { k_tma<16><<<148, 256, 16 * 8192 + 512>>>(src, out, per_cta); }, "TMA bulk ring 16 x 8 KB");
  return 0;
}
