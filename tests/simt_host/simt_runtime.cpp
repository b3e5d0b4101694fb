// TEST INFRASTRUCTURE — runtime half of tests/simt_host/cuda_runtime.h (block scheduler, thread-local indices, error string).
#include <cuda_runtime.h>
#include <stdarg.h>

thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
namespace simt {
thread_local Block* cur = nullptr;
thread_local unsigned linear_tid = 0;

void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()>& body) {
  const unsigned n = block.x * block.y * block.z;
  if (n % 32 != 0) {
    fprintf(stderr, "simt_host: block of %u threads is not a multiple of 32\n", n);
    abort();
  }
  Block blk;
  pthread_barrier_init(&blk.all, nullptr, n);
  blk.warp.resize(n / 32);
  for (auto& w : blk.warp) pthread_barrier_init(&w, nullptr, 32);
  blk.scratch.assign(n, 0ull);
  std::vector<unsigned char> smem(dyn_smem_bytes + 1024, 0xCD);       // garbage-filled like real shared memory
  blk.dyn_smem = smem.data() + ((1024 - (uintptr_t)smem.data() % 1024) % 1024);
  // one OS thread per CUDA thread of a block, created once per launch; the blocks of the grid run one after another on this
  // team (a barrier between blocks: static / dynamic shared memory is reused)
  std::vector<std::thread> threads(n);
  for (unsigned t = 0; t < n; ++t)
    threads[t] = std::thread([&, t]() {
      cur = &blk;
      linear_tid = t;
      threadIdx = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
      blockDim = block;
      gridDim = grid;
      for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
          for (unsigned bx = 0; bx < grid.x; ++bx) {
            blockIdx = uint3{bx, by, bz};
            body();
            pthread_barrier_wait(&blk.all);
          }
    });
  for (auto& th : threads) th.join();
  pthread_barrier_destroy(&blk.all);
  for (auto& w : blk.warp) pthread_barrier_destroy(&w);
}
}  // namespace simt

static thread_local char g_err[512];
__attribute__((weak)) void nsac_set_error(const char* fmt, ...) {   // dense.cu brings its own when it is part of the build
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
extern "C" const char* simt_last_error() { return g_err; }
