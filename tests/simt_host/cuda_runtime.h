// TEST INFRASTRUCTURE — a stand-in for <cuda_runtime.h> that lets g++ compile the plain-SIMT kernels of nopesac_b200/csrc
// (no TMA / tcgen05 / inline PTX) and run them on the host: one OS thread per CUDA thread, blocks one after another,
// __syncthreads = a barrier over the block, warp intrinsics = a barrier over the 32 threads of a warp plus a scratch line,
// atomics = GCC __atomic builtins.  `tests/simt_host/run.py` rewrites `k<<<grid, block, smem, stream>>>(args)` into
// `SIMT_LAUNCH(grid, block, smem, k(args))` and `extern __shared__ T x[];` into a pointer to the block's buffer.  The point: the container has no GPU, so the kernel SOURCE is executed here against the
// oracle before a GPU ever sees it (tests/test_simt_host_planes.py).  Floating point: compiled with -ffp-contract=off so
// that only the explicit __fmaf_rn calls fuse; expf is glibc's (<= 1 ulp, like CUDA's, but not bit-identical).
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(8))) uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(8))) float2 { float x, y; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "simt_host"; }
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t) {
  memcpy(dst, src, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                                     cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < height; ++r) memcpy(static_cast<char*>(dst) + r * dpitch, static_cast<const char*>(src) + r * spitch, width);
  return cudaSuccess;
}

namespace simt {
struct Block {
  pthread_barrier_t all;
  std::vector<pthread_barrier_t> warp;
  std::vector<unsigned long long> scratch;   // 32 words per warp
  unsigned char* dyn_smem;                   // `extern __shared__` storage of the running block (16-byte aligned)
};
extern thread_local Block* cur;
extern thread_local unsigned linear_tid;
}  // namespace simt

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

inline void __syncthreads() { pthread_barrier_wait(&simt::cur->all); }

namespace simt {
// every lane publishes one word, all lanes read the 32 words
inline void exchange(unsigned long long mine, unsigned long long out[32]) {
  const unsigned w = linear_tid >> 5, l = linear_tid & 31;
  cur->scratch[w * 32 + l] = mine;
  pthread_barrier_wait(&cur->warp[w]);
  for (int i = 0; i < 32; ++i) out[i] = cur->scratch[w * 32 + i];
  pthread_barrier_wait(&cur->warp[w]);
}
}  // namespace simt

inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned long long v[32];
  simt::exchange(pred ? 1ull : 0ull, v);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= (unsigned)(v[i] & 1ull) << i;
  return r;
}
inline unsigned __reduce_add_sync(unsigned, unsigned x) {
  unsigned long long v[32];
  simt::exchange(x, v);
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r += (unsigned)v[i];
  return r;
}
inline int __reduce_min_sync(unsigned, int x) {
  unsigned long long v[32];
  simt::exchange((unsigned long long)(unsigned)x, v);
  int r = (int)(unsigned)v[0];
  for (int i = 1; i < 32; ++i) r = (int)(unsigned)v[i] < r ? (int)(unsigned)v[i] : r;
  return r;
}
inline int __reduce_max_sync(unsigned, int x) {
  unsigned long long v[32];
  simt::exchange((unsigned long long)(unsigned)x, v);
  int r = (int)(unsigned)v[0];
  for (int i = 1; i < 32; ++i) r = (int)(unsigned)v[i] > r ? (int)(unsigned)v[i] : r;
  return r;
}
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }
inline float __shfl_xor_sync(unsigned, float x, int o) {
  unsigned long long v[32];
  unsigned bits;
  memcpy(&bits, &x, 4);
  simt::exchange(bits, v);
  bits = (unsigned)v[(simt::linear_tid & 31) ^ o];
  memcpy(&x, &bits, 4);
  return x;
}
inline float __shfl_sync(unsigned, float x, int src) {
  unsigned long long v[32];
  unsigned bits;
  memcpy(&bits, &x, 4);
  simt::exchange(bits, v);
  bits = (unsigned)v[src & 31];
  memcpy(&x, &bits, 4);
  return x;
}
inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&simt::cur->warp[simt::linear_tid >> 5]); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }

template <typename T> inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
template <typename T> inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <typename T> inline int cudaGetSymbolAddress(void** out, T& symbol) { *out = &symbol; return 0; }
template <typename T> inline T atomicMin(T* p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
template <typename T> inline T atomicMax(T* p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (v > old && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}

template <typename T> inline T __ldg(const T* p) { return *p; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __double2float_rn(double d) { return (float)d; }
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline float __expf(float x) { return expf(x); }
inline float __fdividef(float a, float b) { return a / b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }

namespace simt {
void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()>& body);
inline unsigned char* dyn_smem() { return cur->dyn_smem; }
}
#define SIMT_LAUNCH(grid, block, smem, call) simt::launch((grid), (block), (size_t)(smem), [&]() { call; })
