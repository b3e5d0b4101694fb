// Split-bf16 tensor-core GEMM for sm_100a: the dense-contraction engine of the path
//   (hypothesis-generation MLP chain K7, GNN projections K4, pixel-pose convolutions K1 via im2col).
//
//   out[M,N] = act( A[M,K] . W[N,K]^T + bias ),  A = a_hi + a_lo, W = w_hi + w_lo   (16-bit planes)
//            = a_hi.w_hi + a_lo.w_hi + a_hi.w_lo (+ a_lo.w_lo)   accumulated in fp32 (TMEM)
//
// The reference computes these layers in fp32 and the parity bar is 1e-4 abs after ~28 chained layers;
// single-pass bf16 / tf32 misses it.  Two plane formats (`fmt`), same kernel:
//   NSAC_SPLIT_F16  (default) hi = fp16(x), lo = fp16(x - hi): 22 significant bits, ~2^-22 relative per
//                   operand with 3 passes.  Needs |x| <= 65504 (an overflow becomes inf and then NaN poses —
//                   loud, like the reference's own NaN guards); elements below 6e-5 carry <= 3e-8 absolute
//                   error (fp16 subnormals); weight matrices are pre-scaled by a power of two (undone exactly
//                   in the epilogue via `out_scale`) so small weights keep their relative precision.
//   NSAC_SPLIT_BF16 hi = bf16(x), lo = bf16(x - hi): fp32's exponent range but only 16 significant bits —
//                   measured here on the pose chain: 1.0-1.4e-4 on per-hypothesis poses, i.e. over the bar.
// (kind::f16 does not accept mixed fp16/bf16 operands in one MMA — tried: illegal instruction.)
// `passes` = 1 / 2 / 3 / 4 selects hi.hi / + lo.hi / + hi.lo / + lo.lo.
//
// Accumulation accuracy.  The tensor core adds into its fp32 accumulator with truncation, so the error of one
// long tcgen05 accumulation chain grows LINEARLY with the number of MMAs (measured: 3e-9 * K relative, 4e-6 at
// K = 1280 — as large as the bf16-plane representation error).  Two counter-measures, both free of extra
// tensor work: (1) the hi.hi products and the (2^-11 smaller) lo terms go to SEPARATE TMEM accumulators, so
// the lo terms neither lengthen the main chain nor lose bits against it; (2) the K loop is cut into chunks of
// CHUNK_KB x 64 elements that ping-pong between two TMEM buffers while the otherwise idle epilogue warps add
// the finished chunk into fp32 REGISTER accumulators (exact IEEE adds) — the same ping-pong overlaps the
// epilogue of one tile with the MMAs of the next.
//
// Design (one persistent CTA per SM, warp-specialised):
//   warp 0      TMA producer   cp.async.bulk.tensor (SWIZZLE_128B) of the four operand planes into a
//                              3-stage shared-memory ring, mbarrier complete_tx signalling
//   warp 1      MMA issuer     one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BLOCK_N,
//                              K=16) on K-major SW128 descriptors; tcgen05.commit frees ring slots and
//                              publishes the accumulator; also owns tcgen05.alloc / dealloc
//   warps 2..9  epilogue       tcgen05.ld (32x32b.x32) of the fp32 accumulator from TMEM (double-buffered so
//                              the epilogue of tile i overlaps the MMAs of tile i+1), bias (+ per-pair bias
//                              rows), activation, then fp32 rows and/or re-split fp16 / bf16 hi/lo planes for the
//                              next layer straight from registers.  Two warps per TMEM lane quadrant (each owns
//                              half of the columns), packed fp32 math, vector bias loads: one 128x128 tile takes
//                              3.2 us instead of 10.6 us (nsac_debug_gemm_trace) - the epilogue, not the tensor
//                              pipe, bounded the engine for K <= ~1300 and put a 12 us floor under every launch
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                 // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;                // two per TMEM lane quadrant: each owns half of the tile's columns
constexpr int NUM_THREADS = 64 + EPI_WARPS * 32;   // TMA warp + MMA warp + epilogue warps
constexpr uint32_t SPIN_LIMIT = 1u << 24;  // suspended waits of up to ~1 us each: ~17 s, then trap instead of hanging

constexpr int CHUNK_KB = 4;                 // K-blocks (x64 elements) accumulated inside the tensor core per chunk

template <int BLOCK_N, int RES = 0>      // RES: 0 = plain epilogue; residual epilogue with 1 = 3 operand stages + 2 tile buffers, 2 = 2 + 4
struct Cfg {
  // RES (residual epilogue, 64-wide tiles only): two tile buffers [128 x 64] x (hi, lo) = 2 x 32 KB.  The residual tile is
  // TMA-prefetched into one of them while the tile's MMAs run, the epilogue turns it IN PLACE into the output tile, and a TMA
  // store writes it out asynchronously - no per-warp staging buffers, no latency-exposed global loads in the epilogue.
  // RES = 3: residual read straight from global memory in the plain row-per-lane epilogue of the 128-wide kernel (no tile
  // buffers, no TMA store): for the deep-K layers (res4 / res5), where 64-wide tiles are bound by re-fetching A from L2
  static constexpr bool RES_TMA = RES == 1 || RES == 2;
  static_assert(!RES_TMA || BLOCK_N == 64, "the in-place residual epilogue uses 64-wide tiles");
  static_assert(RES != 3 || BLOCK_N == 128, "the global-residual epilogue is the 128-wide kernel's");
  static constexpr int STAGES = RES == 2 ? 2 : (RES == 1 ? 3 : (BLOCK_N == 64 ? 4 : 3));       // 48 KB / 64 KB per stage
  static constexpr int RES_BUFS = RES == 2 ? 4 : (RES == 1 ? 2 : 0);    // tile buffers: residual prefetch that many tiles ahead, stores drain behind
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;     // one plane
  static constexpr int W_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
  static constexpr int ACC_COLS = 2 * BLOCK_N;              // one buffer = hi.hi accumulator + lo-terms accumulator
  static constexpr int TMEM_COLS = 2 * ACC_COLS;            // two buffers (ping-pong between chunks / tiles)
  static constexpr int RES_PLANE_BYTES = RES_TMA ? BLOCK_M * 128 : 0;          // [128 rows x 64 columns] of one plane, 128-byte swizzled rows
  static constexpr int TILE_BUF_BYTES = 2 * RES_PLANE_BYTES;               // hi + lo
  static constexpr int RES_BYTES = RES_BUFS * TILE_BUF_BYTES;
  static constexpr int OUT_STAGE_BYTES = RES_TMA ? 0 : EPI_WARPS * 4096;  // per epilogue warp: 32 rows x 128 B, to turn row-per-lane data into coalesced stores
  static constexpr int THREADS = NUM_THREADS + (RES_TMA ? 32 : 0);             // RES: + one warp that owns the TMA stores of the output tiles
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + RES_BYTES + 1024 /*align*/ + 256 /*barriers*/ + OUT_STAGE_BYTES;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(TMEM_COLS <= 512, "TMEM has 512 columns");
};

// ------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Waiting warps must not steal issue slots from the working warps: try_wait with a suspend-time hint parks the
// warp in hardware until the phase completes (the first ncu capture of the scoring kernel had 40 % of all issued
// instructions in plain try_wait spin loops).  Bounded: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity), "r"(1000u) : "memory");
    if (done) break;
    if (++spins > SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* smem_src, const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (=1, unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 (= 8 rows x 128 B = 1024 -> 64) | [46,48) version = 1 | [61,64) layout = 2
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b format (F16 = 0, BF16 = 1) @7/@10, K-major both,
// N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool bf16) {
  return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

// Optional timeline of CTA 0 (profiling aid, nsac_debug_gemm_trace): %globaltimer at the role hand-offs of the launch.
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define GEMM_TRACE(slot)                                                        \
  do {                                                                          \
    if (p.trace && blockIdx.x == 0) p.trace[slot] = gtime();                    \
  } while (0)

struct GemmParams {
  unsigned long long* trace;   // nullptr unless tracing: [16] slots
  const float* bias;
  int bias_group_rows;
  int M, N, K, act, passes, fmt;
  // implicit-GEMM 3x3 / stride 1 / pad 1 convolution over NHWC planes (conv_taps == 9), else plain GEMM
  int conv_taps, H, W, BW, BH, cblocks, tiles_x, tiles_y;      // H, W: OUTPUT map; conv_taps: 9 = 3x3 (pad 1), 1 = 1x1 (conv mode only)
  int conv;                 // 1 = implicit-GEMM convolution (A gathered from NHWC planes by 4-D TMA), 0 = plain GEMM
  int cstride;              // convolution stride (1 or 2): tap (dy, dx) of output pixel (y, x) reads input (cstride*y + dy, cstride*x + dx)
  float out_scale;          // multiplies the accumulator before bias (undoes the power-of-two weight pre-scale)
  float* out_f32;
  int ldo;
  uint16_t* out_hi;
  uint16_t* out_lo;
  int ld_split;
  // optional residual input (16-bit hi/lo planes, same format as the operands): out = act(acc * scale + bias + (res_hi + res_lo))
  const uint16_t* res_hi;
  const uint16_t* res_lo;
  int ld_res;
  const float* row_bias;    // optional per-ROW scalar added with the bias (out[m, n] += row_bias[m]): the mask-logit GEMM of
                            // PlaneTRHead, whose M rows are plane queries (planeTR_head.py:148-150)
  int a_lo_zero;            // the A operand has no lo plane (exactly representable in 16 bits, e.g. raw 8-bit pixels): the
                            // lo.hi pass and its TMA loads are skipped, passes = 3 then means hi.hi + hi.lo
};

// Sticky overflow flag (nsac_plane_overflow): set when a FINITE value that does not fit an fp16 plane (|x| > 65504) is written as
// a plane by nsac_split16 or a GEMM epilogue.  Overflow used to be "loud" only through the inf -> NaN it usually causes downstream;
// the r2k robustness test found inputs where normalisation layers turned it back into finite, plausible, WRONG poses.
__device__ unsigned int g_plane_overflow = 0u;
__device__ __forceinline__ void note_overflow(float maxabs) {
  if (maxabs > 65504.f && maxabs < INFINITY) atomicOr(&g_plane_overflow, 1u);
}

// x -> (hi, lo) 16-bit planes in the requested format
__device__ __forceinline__ void split16(float x, int fmt, uint16_t& hi, uint16_t& lo) {
  if (fmt == NSAC_SPLIT_F16) {
    note_overflow(fabsf(x));
    const __half h = __float2half_rn(x);
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(__float2half_rn(x - __half2float(h)));
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(__float2bfloat16_rn(x - __bfloat162float(h)));
  }
}




// ---- packed fp32 (sm_100 FADD2 / FFMA2) for the epilogue: it is a long serial instruction stream on one warp per
// scheduler, and for K <= ~1300 the tile's MMAs finish before it does (nsac_debug_gemm_trace: 8-10 us per 128x128 tile
// before this rewrite, i.e. the epilogue - not the tensor pipe - bounded the engine)
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// two values -> packed (hi, lo) 16-bit plane words
template <int FMT>
__device__ __forceinline__ void split16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  if (FMT == NSAC_SPLIT_F16) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 back = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - back.x, b - back.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  } else {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 back = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(a - back.x, b - back.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
  }
}

struct EpiOut {
  float out_scale;
  float slope;            // activation as max(x, slope * x): 1 = none, 0 = ReLU, 0.01 = LeakyReLU (one code path: the epilogue is
                          // straight-line unrolled code and six ACT x FMT copies of it thrashed the instruction cache)
  const float* brow;      // bias row of this thread's output row (nullptr: no bias)
  float* out_f32;         // row pointers (nullptr: not requested)
  uint16_t* out_hi;
  uint16_t* out_lo;
  int n_valid;            // valid columns from col0 on (>= 32: full chunk)
  bool vec_ok;            // 16-byte aligned rows: vector loads / stores allowed
  float row_bias;         // per-row scalar added with the bias (0 if none)
  const uint16_t* res_hi; // RES = 3: this row's residual planes (nullptr: none); added before the activation
  const uint16_t* res_lo;
};

// packed 16-bit plane word -> two floats
template <int FMT>
__device__ __forceinline__ float2 unsplit16x2(uint32_t w) {
  if (FMT == NSAC_SPLIT_F16) return __half22float2(*reinterpret_cast<const __half2*>(&w));
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}

// 32 accumulator columns -> scale, bias, activation, fp32 store, split-plane store
template <int FMT>
__device__ __forceinline__ void finish32(const u64 (&acc)[16], int col0, const EpiOut& o) {
  float f[32];
  const u64 sc = pk2(o.out_scale, o.out_scale);
  const bool full = o.n_valid >= 32 && o.vec_ok;
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    u64 b01 = 0ull, b23 = 0ull;
    if (o.brow) {
      if (full) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(o.brow + col0 + i));
        b01 = pk2(b4.x, b4.y); b23 = pk2(b4.z, b4.w);
      } else {
        float b[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) b[k] = i + k < o.n_valid ? __ldg(o.brow + col0 + i + k) : 0.f;
        b01 = pk2(b[0], b[1]); b23 = pk2(b[2], b[3]);
      }
    }
    const u64 rb = pk2(o.row_bias, o.row_bias);
    upk2(ffma2(acc[i >> 1], sc, fadd2(b01, rb)), f[i], f[i + 1]);
    upk2(ffma2(acc[(i >> 1) + 1], sc, fadd2(b23, rb)), f[i + 2], f[i + 3]);
  }
  if (o.res_hi) {          // slow path of the global-residual epilogue (ragged tiles): element loads
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (i < o.n_valid) {
        const uint32_t h = o.res_hi[col0 + i], l = o.res_lo[col0 + i];
        f[i] += unsplit16x2<FMT>(h).x + unsplit16x2<FMT>(l).x;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], fmaf(o.slope, f[i], 0.f));      // (+0 addend: ReLU of a negative is +0, not -0)
  if (o.out_f32) {
    float* dst = o.out_f32 + col0;
    if (full) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i < o.n_valid) dst[i] = f[i];
    }
  }
  if (o.out_hi) {
    uint32_t hi[16], lo[16];
    if (FMT == NSAC_SPLIT_F16) {
      float mx = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, i < o.n_valid ? fabsf(f[i]) : 0.f);
      note_overflow(mx);
    }
#pragma unroll
    for (int i = 0; i < 32; i += 2) split16x2<FMT>(f[i], f[i + 1], hi[i >> 1], lo[i >> 1]);
    uint16_t* dh = o.out_hi + col0;
    uint16_t* dl = o.out_lo + col0;
    if (full) {
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        *reinterpret_cast<uint4*>(dh + 2 * i) = make_uint4(hi[i], hi[i + 1], hi[i + 2], hi[i + 3]);
        *reinterpret_cast<uint4*>(dl + 2 * i) = make_uint4(lo[i], lo[i + 1], lo[i + 2], lo[i + 3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (i < o.n_valid) {
          dh[i] = (i & 1) ? (uint16_t)(hi[i >> 1] >> 16) : (uint16_t)(hi[i >> 1] & 0xffff);
          dl[i] = (i & 1) ? (uint16_t)(lo[i >> 1] >> 16) : (uint16_t)(lo[i >> 1] & 0xffff);
        }
      }
    }
  }
}


// Row-per-lane data -> coalesced global stores.  After the accumulator read every lane holds one output ROW; written
// straight from registers a 16-byte store instruction touches 32 different 128-byte lines (the LSU then needs ~32 cycles
// per instruction: 2.8 of the 3.2 us of a tile's epilogue, nsac_debug_gemm_trace).  Each warp therefore transposes
// 32 rows x 128 (or 64) bytes through its 4 KB of shared memory (16-byte chunks XOR-swizzled by the row: conflict-free
// both ways) and writes them out as 4 full lines (8 half lines) per instruction.  row_ptr = this lane's own destination
// (start of its row segment).
template <int WORDS>      // WORDS = 32: 128-byte row segments, WORDS = 16: 64-byte row segments
__device__ __forceinline__ void staged_store(uint8_t* stage, const uint32_t (&w)[WORDS], uint8_t* row_ptr, int lane) {
  constexpr int BYTES = WORDS * 4, CPR = BYTES / 16, RPI = 32 / CPR;     // 16-byte chunks per row, rows per store instruction
  const int swz_w = CPR == 8 ? (lane & 7) : ((lane >> 1) & 3);
#pragma unroll
  for (int c = 0; c < CPR; ++c)
    *reinterpret_cast<uint4*>(stage + lane * BYTES + ((c ^ swz_w) << 4)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
  __syncwarp();
  const int sub = lane / CPR, ch = lane % CPR;
  const unsigned long long my = reinterpret_cast<unsigned long long>(row_ptr);
#pragma unroll
  for (int j = 0; j < 32 / RPI; ++j) {
    const int row = RPI * j + sub;
    const int swz_r = CPR == 8 ? (row & 7) : ((row >> 1) & 3);
    const uint4 x = *reinterpret_cast<const uint4*>(stage + row * BYTES + ((ch ^ swz_r) << 4));
    uint8_t* dst = reinterpret_cast<uint8_t*>(__shfl_sync(0xffffffffu, my, row));
    *reinterpret_cast<uint4*>(dst + ch * 16) = x;
  }
  __syncwarp();
}

// The reverse: coalesced global loads -> row-per-lane data.  row_ptr = this lane's own source (start of its row segment of
// WORDS * 4 bytes); every load instruction reads full lines (the rows of `RPI` lanes at a time) into the warp's staging buffer,
// then each lane picks up its own row.
template <int WORDS>
__device__ __forceinline__ void staged_load(uint8_t* stage, uint32_t (&w)[WORDS], const uint8_t* row_ptr, int lane) {
  constexpr int BYTES = WORDS * 4, CPR = BYTES / 16, RPI = 32 / CPR;
  const int sub = lane / CPR, ch = lane % CPR;
  const unsigned long long my = reinterpret_cast<unsigned long long>(row_ptr);
#pragma unroll
  for (int j = 0; j < 32 / RPI; ++j) {
    const int row = RPI * j + sub;
    const int swz_r = CPR == 8 ? (row & 7) : ((row >> 1) & 3);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(__shfl_sync(0xffffffffu, my, row));
    *reinterpret_cast<uint4*>(stage + row * BYTES + ((ch ^ swz_r) << 4)) = __ldg(reinterpret_cast<const uint4*>(src + ch * 16));
  }
  __syncwarp();
  const int swz_w = CPR == 8 ? (lane & 7) : ((lane >> 1) & 3);
#pragma unroll
  for (int c = 0; c < CPR; ++c) {
    const uint4 x = *reinterpret_cast<const uint4*>(stage + lane * BYTES + ((c ^ swz_w) << 4));
    w[4 * c] = x.x; w[4 * c + 1] = x.y; w[4 * c + 2] = x.z; w[4 * c + 3] = x.w;
  }
  __syncwarp();
}

// 64 accumulator columns of a full tile part (every lane of the warp has a valid row, all 64 columns valid, aligned):
// scale, bias, activation, then fp32 rows and / or split planes through the staged stores, 32 columns at a time
template <int FMT, int GROUPS>
__device__ __forceinline__ void finish64_staged(const u64 (&sum)[16 * GROUPS], int col0, const EpiOut& o, uint8_t* stage, int lane) {
  const u64 sc = pk2(o.out_scale, o.out_scale), sl = pk2(o.slope, o.slope);
#pragma unroll
  for (int g = 0; g < GROUPS; ++g) {       // groups of 32 columns (two for 128-wide tiles, one for 64-wide tiles)
    uint32_t f[32];                        // fp32 bit patterns
    uint32_t rh[16], rl[16];               // RES = 3: this row's 32 residual columns, both planes (2 x 64 contiguous bytes),
    if (o.res_hi) {                        // fetched with full-line loads through the staging buffer (warp-uniform branch)
      staged_load<16>(stage, rh, reinterpret_cast<const uint8_t*>(o.res_hi + col0 + 32 * g), lane);
      staged_load<16>(stage, rl, reinterpret_cast<const uint8_t*>(o.res_lo + col0 + 32 * g), lane);
    }
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      u64 b01 = 0ull, b23 = 0ull;
      if (o.brow) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(o.brow + col0 + 32 * g + i));
        b01 = pk2(b4.x, b4.y); b23 = pk2(b4.z, b4.w);
      }
      if (o.res_hi) {                      // bias + residual (hi + lo reproduces the fp32 activation)
        const uint32_t hw0 = rh[i >> 1], hw1 = rh[(i >> 1) + 1], lw0 = rl[i >> 1], lw1 = rl[(i >> 1) + 1];
        const float2 a0 = unsplit16x2<FMT>(hw0), a1 = unsplit16x2<FMT>(hw1), c0 = unsplit16x2<FMT>(lw0), c1 = unsplit16x2<FMT>(lw1);
        b01 = fadd2(b01, fadd2(pk2(a0.x, a0.y), pk2(c0.x, c0.y)));
        b23 = fadd2(b23, fadd2(pk2(a1.x, a1.y), pk2(c1.x, c1.y)));
      }
      float x0, x1, x2, x3;
      const u64 rb = pk2(o.row_bias, o.row_bias);
      upk2(ffma2(sum[16 * g + (i >> 1)], sc, fadd2(b01, rb)), x0, x1);
      upk2(ffma2(sum[16 * g + (i >> 1) + 1], sc, fadd2(b23, rb)), x2, x3);
      {   // activation as max(x, slope * x + 0): the +0 addend makes ReLU of a negative +0, not -0
        float t0, t1, t2, t3;
        upk2(ffma2(pk2(x0, x1), sl, 0ull), t0, t1);
        upk2(ffma2(pk2(x2, x3), sl, 0ull), t2, t3);
        x0 = fmaxf(x0, t0); x1 = fmaxf(x1, t1); x2 = fmaxf(x2, t2); x3 = fmaxf(x3, t3);
      }
      f[i] = __float_as_uint(x0); f[i + 1] = __float_as_uint(x1); f[i + 2] = __float_as_uint(x2); f[i + 3] = __float_as_uint(x3);
    }
    if (o.out_f32)                         // 32 fp32 columns = one 128-byte segment per row
      staged_store<32>(stage, f, reinterpret_cast<uint8_t*>(o.out_f32 + col0 + 32 * g), lane);
    if (o.out_hi) {                        // 32 16-bit columns = one 64-byte segment per row and plane
      uint32_t hi[16], lo[16];
      if (FMT == NSAC_SPLIT_F16) {           // one FMNMX per element on the ALU pipe; the atomic only fires on an overflow
        float mx = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fabsf(__uint_as_float(f[i])));
        note_overflow(mx);
      }
#pragma unroll
      for (int i = 0; i < 32; i += 2) split16x2<FMT>(__uint_as_float(f[i]), __uint_as_float(f[i + 1]), hi[i >> 1], lo[i >> 1]);
      staged_store<16>(stage, hi, reinterpret_cast<uint8_t*>(o.out_hi + col0 + 32 * g), lane);
      staged_store<16>(stage, lo, reinterpret_cast<uint8_t*>(o.out_lo + col0 + 32 * g), lane);
    }
  }
}

template <int BLOCK_N, int FMT, int RES>
__global__ void __launch_bounds__((Cfg<BLOCK_N, RES>::THREADS), 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                   const __grid_constant__ CUtensorMap map_r_hi, const __grid_constant__ CUtensorMap map_r_lo,
                   const __grid_constant__ CUtensorMap map_o_hi, const __grid_constant__ CUtensorMap map_o_lo,
                   const GemmParams p) {
  using C = Cfg<BLOCK_N, RES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-window pointer
  uint8_t* res_stage = smem + C::STAGES * C::STAGE_BYTES;            // RES: two tile buffers [hi | lo][128 rows x 128 B], swizzled
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES + C::RES_BYTES);
  uint64_t* full = bars;                       // [STAGES]
  uint64_t* empty = bars + C::STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * C::STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  constexpr int RB = C::RES_BUFS > 0 ? C::RES_BUFS : 1;
  uint64_t* res_full = tmem_empty + 2;         // [RB]
  uint64_t* res_empty = res_full + RB;         // [RB]
  uint64_t* tile_ready = res_empty + RB;       // [RB]  RES: the epilogue warps have turned tile buffer b into the output tile
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tile_ready + RB);
  uint8_t* out_stage = smem + C::STAGES * C::STAGE_BYTES + C::RES_BYTES + 256;      // [EPI_WARPS][4 KB]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool conv = p.conv != 0;
  const int tiles_per_img = p.tiles_x * p.tiles_y;     // conv: an M tile is a BW x BH pixel patch of one image
  const int tiles_m = conv ? (p.M / (p.H * p.W)) * tiles_per_img : (p.M + BLOCK_M - 1) / BLOCK_M;
  const int tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = p.K / BLOCK_K;

  if (threadIdx.x == 0) GEMM_TRACE(0);
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w_hi)) : "memory");
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], EPI_WARPS); }
    for (int a = 0; a < RB; ++a) { mbar_init(&res_full[a], 1); mbar_init(&res_empty[a], 1); mbar_init(&tile_ready[a], EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM allocation (whole warp), address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"((uint32_t)C::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  if (threadIdx.x == 0) GEMM_TRACE(1);

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0, tile_ctr = 0;
      const uint32_t a_bytes = conv ? (uint32_t)(p.BW * p.BH * BLOCK_K * 2) : (uint32_t)C::A_BYTES;   // box bytes
      const bool load_alo = p.passes >= 2 && !p.a_lo_zero;
      const uint32_t tx = (load_alo ? 2 * a_bytes : a_bytes) + (p.passes >= 3 ? 2 * C::W_BYTES : C::W_BYTES);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        // n fastest: the CTAs running at the same time share a few M blocks, so every A tile comes from HBM once and its other
        // N tiles hit L2 (m-fastest order swept the whole A matrix once per N tile: 630 MB x 4 for the 64 -> 256 layers of res2)
        const int tm = tile / tiles_n;
        const int m0 = tm * BLOCK_M, n0 = (tile % tiles_n) * BLOCK_N;
        int img = 0, y0 = 0, x0 = 0;
        if (conv) {
          img = tm / tiles_per_img;
          const int t2 = tm % tiles_per_img;
          y0 = (t2 / p.tiles_x) * p.BH;
          x0 = (t2 % p.tiles_x) * p.BW;
        }
        if (C::RES_TMA) {     // this tile's residual [128 x 64] x (hi, lo) into a tile buffer: lands while the MMAs run
          const uint32_t b = tile_ctr % RB, use = tile_ctr / RB;
          mbar_wait(&res_empty[b], (use & 1) ^ 1);            // the TMA store of the buffer's previous tile has read it out
          mbar_expect_tx(&res_full[b], (uint32_t)C::TILE_BUF_BYTES);
          uint8_t* tb = res_stage + b * C::TILE_BUF_BYTES;
          tma_load_2d(tb, &map_r_hi, &res_full[b], n0, m0);
          tma_load_2d(tb + C::RES_PLANE_BYTES, &map_r_lo, &res_full[b], n0, m0);
          ++tile_ctr;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * C::STAGE_BYTES;
          mbar_expect_tx(&full[stage], tx);
          if (conv) {
            // k-block = (filter tap, 64-channel block): the A tile is the input patch shifted by the tap; TMA
            // zero-fills the out-of-image halo (padding = 1), so no im2col matrix ever exists
            const int tap = kb / p.cblocks, cb = kb - tap * p.cblocks;
            const int dy = p.conv_taps == 9 ? tap / 3 - 1 : 0, dx = p.conv_taps == 9 ? tap % 3 - 1 : 0;
            const int xi = x0 * p.cstride + dx, yi = y0 * p.cstride + dy;     // (strided maps: the tensor map's traversal stride picks every cstride-th pixel)
            tma_load_4d(st, &map_a_hi, &full[stage], cb * BLOCK_K, xi, yi, img);
            if (load_alo) tma_load_4d(st + C::A_BYTES, &map_a_lo, &full[stage], cb * BLOCK_K, xi, yi, img);
          } else {
            tma_load_2d(st, &map_a_hi, &full[stage], kb * BLOCK_K, m0);
            if (load_alo) tma_load_2d(st + C::A_BYTES, &map_a_lo, &full[stage], kb * BLOCK_K, m0);
          }
          tma_load_2d(st + 2 * C::A_BYTES, &map_w_hi, &full[stage], kb * BLOCK_K, n0);
          if (p.passes >= 3) tma_load_2d(st + 2 * C::A_BYTES + C::W_BYTES, &map_w_lo, &full[stage], kb * BLOCK_K, n0);
          if (tile == (int)blockIdx.x && kb == 0) GEMM_TRACE(2);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, FMT == NSAC_SPLIT_BF16);
      int stage = 0; uint32_t phase = 0;
      uint32_t chunk_ctr = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int kb0 = 0; kb0 < num_kb; kb0 += CHUNK_KB, ++chunk_ctr) {
          const uint32_t buf = chunk_ctr & 1, buf_phase = (chunk_ctr >> 1) & 1;
          mbar_wait(&tmem_empty[buf], buf_phase ^ 1);
          tcgen05_fence_after();
          const uint32_t d_main = tmem_base + buf * C::ACC_COLS, d_lo = d_main + BLOCK_N;
          const int kb1 = kb0 + CHUNK_KB < num_kb ? kb0 + CHUNK_KB : num_kb;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full[stage], phase);
            if (tile == (int)blockIdx.x && kb == 0) GEMM_TRACE(3);
            tcgen05_fence_after();
            const uint32_t st = smem_u32(smem + stage * C::STAGE_BYTES);
            const uint64_t a_hi = make_sw128_desc(st), a_lo = make_sw128_desc(st + C::A_BYTES);
            const uint64_t w_hi = make_sw128_desc(st + 2 * C::A_BYTES), w_lo = make_sw128_desc(st + 2 * C::A_BYTES + C::W_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              const uint64_t koff = (uint64_t)((k * UMMA_K * 2) >> 4);   // advance the start address inside the swizzle row
              const uint32_t first = (kb == kb0 && k == 0) ? 0u : 1u;
              umma_bf16(d_main, a_hi + koff, w_hi + koff, idesc, first);
              uint32_t lo_first = first;
              if (p.passes >= 2 && !p.a_lo_zero) { umma_bf16(d_lo, a_lo + koff, w_hi + koff, idesc, first); lo_first = 1u; }
              if (p.passes >= 3) umma_bf16(d_lo, a_hi + koff, w_lo + koff, idesc, lo_first);
              if (p.passes >= 4 && !p.a_lo_zero) umma_bf16(d_lo, a_lo + koff, w_lo + koff, idesc, 1);
            }
            tcgen05_commit(&empty[stage]);                 // slot reusable once these MMAs retire
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          }
          tcgen05_commit(&tmem_full[buf]);                 // chunk complete -> epilogue warps
          if (tile == (int)blockIdx.x && kb0 == 0) GEMM_TRACE(4);
        }
      }
    }
  } else if (C::RES_TMA && warp == 2 + EPI_WARPS) {
    // ===================================================================== TMA store warp (residual epilogue only)
    // Takes everything after the epilogue math off the epilogue warps' critical path: they arrive on tile_ready[b] and go on to the
    // next tile; this thread stores the finished tile, waits until the store has READ the buffer and hands it back to the producer.
    if (lane == 0) {
      uint32_t ctr = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ctr) {
        const uint32_t b = ctr % RB;
        mbar_wait(&tile_ready[b], (ctr / RB) & 1);         // (the writers issued fence.proxy.async before arriving)
        uint8_t* tb = res_stage + b * C::TILE_BUF_BYTES;
        const int m0 = (tile / tiles_n) * BLOCK_M, nt = (tile % tiles_n) * BLOCK_N;
        tma_store_2d(tb, &map_o_hi, nt, m0);               // rows >= M / columns >= N are clipped by the tensor map
        tma_store_2d(tb + C::RES_PLANE_BYTES, &map_o_lo, nt, m0);
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(&res_empty[b]);                         // the producer may load the residual of tile t + RES_BUFS into it
      }
      tma_store_wait_all();                                 // the last stores must have left shared memory before the CTA exits
    }
  } else {
    // ===================================================================== epilogue warps 2..9
    // TMEM lane quadrant = warp % 4 (hardware rule); the two warps of a quadrant split the tile's columns in halves
    const int quad = warp & 3, half = (warp - 2) >> 2;
    constexpr int HALF_N = BLOCK_N / 2;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    uint32_t chunk_ctr = 0, res_ctr = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * BLOCK_M, n0 = (tile % tiles_n) * BLOCK_N + half * HALF_N;
      u64 sum[HALF_N / 2];
#pragma unroll
      for (int i = 0; i < HALF_N / 2; ++i) sum[i] = 0ull;
      for (int kb0 = 0; kb0 < num_kb; kb0 += CHUNK_KB, ++chunk_ctr) {
        const uint32_t buf = chunk_ctr & 1, buf_phase = (chunk_ctr >> 1) & 1;
        mbar_wait(&tmem_full[buf], buf_phase);
        if (tile == (int)blockIdx.x && kb0 == 0 && threadIdx.x == 64) GEMM_TRACE(5);
        tcgen05_fence_after();
        const uint32_t t_main = tmem_base + lane_base + buf * C::ACC_COLS + half * HALF_N;
#pragma unroll
        for (int c = 0; c < HALF_N; c += 32) {
          uint32_t v[32];
          tmem_ld32(t_main + c, v);
          if (p.passes >= 3 || (p.passes >= 2 && !p.a_lo_zero)) {
            uint32_t u[32];
            tmem_ld32(t_main + BLOCK_N + c, u);
#pragma unroll
            for (int i = 0; i < 32; i += 2)
              sum[(c + i) >> 1] = fadd2(sum[(c + i) >> 1], fadd2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])),
                                                                 pk2(__uint_as_float(u[i]), __uint_as_float(u[i + 1]))));
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2)
              sum[(c + i) >> 1] = fadd2(sum[(c + i) >> 1], pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[buf]);
      }
      if (tile == (int)blockIdx.x && threadIdx.x == 64) GEMM_TRACE(9);
      if constexpr (C::RES_TMA) {
        // ---- residual epilogue: out = act(acc * scale + bias + residual), IN PLACE in the tile buffer, then one TMA store.
        // lane = tile row quad * 32 + lane, this warp's 32 columns = 16-byte chunks half * 4 .. half * 4 + 3 of the 128-byte row
        const uint32_t b = res_ctr % RB;
        mbar_wait(&res_full[b], (res_ctr / RB) & 1);
        uint8_t* tb = res_stage + b * C::TILE_BUF_BYTES;
        uint8_t* rowp = tb + (quad * 32 + lane) * 128;
        const u64 sc = pk2(p.out_scale, p.out_scale);
        const float slope = p.act == NSAC_ACT_RELU ? 0.f : (p.act == NSAC_ACT_LEAKY ? 0.01f : 1.f);
        const u64 sl = pk2(slope, slope);
        float ovf = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int off = ((half * 4 + c) ^ (lane & 7)) << 4;
          const uint4 rh = *reinterpret_cast<const uint4*>(rowp + off);
          const uint4 rl = *reinterpret_cast<const uint4*>(rowp + C::RES_PLANE_BYTES + off);
          const uint32_t rhw[4] = {rh.x, rh.y, rh.z, rh.w}, rlw[4] = {rl.x, rl.y, rl.z, rl.w};
          uint32_t oh[4], ol[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {                  // two columns per word
            const int col = n0 + 8 * c + 2 * j;
            u64 bb = 0ull;
            if (p.bias) bb = pk2(col < p.N ? __ldg(p.bias + col) : 0.f, col + 1 < p.N ? __ldg(p.bias + col + 1) : 0.f);
            const float2 h = unsplit16x2<FMT>(rhw[j]), l = unsplit16x2<FMT>(rlw[j]);
            bb = fadd2(bb, fadd2(pk2(h.x, h.y), pk2(l.x, l.y)));      // bias + residual (hi + lo reproduces the fp32 activation)
            float x0, x1, t0, t1;
            upk2(ffma2(sum[4 * c + j], sc, bb), x0, x1);
            upk2(ffma2(pk2(x0, x1), sl, 0ull), t0, t1);               // activation as max(x, slope * x + 0)
            const float y0v = fmaxf(x0, t0), y1v = fmaxf(x1, t1);
            ovf = fmaxf(ovf, fmaxf(fabsf(y0v), fabsf(y1v)));
            split16x2<FMT>(y0v, y1v, oh[j], ol[j]);
          }
          *reinterpret_cast<uint4*>(rowp + off) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
          *reinterpret_cast<uint4*>(rowp + C::RES_PLANE_BYTES + off) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
        }
        if (FMT == NSAC_SPLIT_F16 && m0 + quad * 32 + lane < p.M) note_overflow(ovf);
        fence_proxy_async_smem();                        // generic-proxy writes -> visible to the TMA (async proxy) store
        __syncwarp();
        if (lane == 0) mbar_arrive(&tile_ready[b]);      // 8 arrivals = the whole tile is written: the store warp takes over
        ++res_ctr;
        continue;
      }
      // ---- bias, activation, stores
      int row = m0 + quad * 32 + lane;
      bool row_ok = row < p.M;
      if (conv) {      // tile row r = pixel (y0 + r / BW, x0 + r % BW) of image tm / tiles_per_img
        const int tm = tile / tiles_n, img = tm / tiles_per_img, t2 = tm % tiles_per_img;
        const int r = quad * 32 + lane;
        const int y = (t2 / p.tiles_x) * p.BH + r / p.BW, x = (t2 % p.tiles_x) * p.BW + r % p.BW;
        row_ok = r < p.BW * p.BH && y < p.H && x < p.W;
        row = (img * p.H + y) * p.W + x;
      }
      EpiOut o;
      o.out_scale = p.out_scale;
      o.slope = p.act == NSAC_ACT_RELU ? 0.f : (p.act == NSAC_ACT_LEAKY ? 0.01f : 1.f);
      o.brow = nullptr; o.out_f32 = nullptr; o.out_hi = nullptr; o.out_lo = nullptr; o.vec_ok = false; o.n_valid = 0;
      o.res_hi = nullptr; o.res_lo = nullptr;
      o.row_bias = (p.row_bias && row_ok) ? __ldg(p.row_bias + row) : 0.f;
      if (row_ok) {
        if (p.bias) o.brow = p.bias_group_rows > 0 ? p.bias + (size_t)(row / p.bias_group_rows) * p.N : p.bias;
        o.out_f32 = p.out_f32 ? p.out_f32 + (size_t)row * p.ldo : nullptr;
        o.out_hi = p.out_hi ? p.out_hi + (size_t)row * p.ld_split : nullptr;
        o.out_lo = p.out_lo ? p.out_lo + (size_t)row * p.ld_split : nullptr;
        if (RES == 3) { o.res_hi = p.res_hi + (size_t)row * p.ld_res; o.res_lo = p.res_lo + (size_t)row * p.ld_res; }
        o.vec_ok = (RES != 3 || ((reinterpret_cast<uintptr_t>(o.res_hi) | reinterpret_cast<uintptr_t>(o.res_lo)) & 15) == 0) &&
                   (!o.brow || (reinterpret_cast<uintptr_t>(o.brow) & 15) == 0) &&
                   (!o.out_f32 || (reinterpret_cast<uintptr_t>(o.out_f32) & 15) == 0) &&
                   (!o.out_hi || ((reinterpret_cast<uintptr_t>(o.out_hi) | reinterpret_cast<uintptr_t>(o.out_lo)) & 15) == 0);
      }
      // fast path (warp-uniform): every lane has a valid, aligned row and all 64 columns of this half exist
      const bool staged = __all_sync(0xffffffffu, row_ok && o.vec_ok) && n0 + HALF_N <= p.N;
      if (staged) {
        uint8_t* stage = out_stage + (warp - 2) * 4096;
        finish64_staged<FMT, HALF_N / 32>(sum, n0, o, stage, lane);
      } else if (row_ok) {
#pragma unroll
        for (int c = 0; c < HALF_N; c += 32) {
          const int col0 = n0 + c;
          if (col0 < p.N) {
            o.n_valid = p.N - col0;
            u64 acc[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = sum[(c >> 1) + i];
            finish32<FMT>(acc, col0, o);
          }
        }
      }
    }
  }
  if (threadIdx.x == 64) GEMM_TRACE(6);
  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) GEMM_TRACE(7);
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS));
  }
  if (threadIdx.x == 32) GEMM_TRACE(8);
}

// fp32 [rows, K] * scale -> hi / lo 16-bit planes [rows, ld_split], zero-padded to ld_split columns
__global__ void split16_kernel(const float* __restrict__ x, int ldx, int rows, int K, float scale, int fmt,
                               uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int ld_split) {
  const size_t total = (size_t)rows * ld_split;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx / ld_split), c = (int)(idx - (size_t)r * ld_split);
    const float v = c < K ? x[(size_t)r * ldx + c] * scale : 0.f;
    uint16_t h, l;
    split16(v, fmt, h, l);
    hi[idx] = h;
    lo[idx] = l;
  }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// row-major bf16 [rows, K] with row stride ld (elements): box = [box_rows x 64 elements], 128-byte swizzle
bool make_map(CUtensorMap* map, const void* base, int rows, int K, int ld, int box_rows, bool is_bf16) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BLOCK_N, int RES>
int launch_gemm(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& wh, const CUtensorMap& wl, const CUtensorMap& rh,
                const CUtensorMap& rl, const CUtensorMap& oh, const CUtensorMap& ol, const GemmParams& p, int tiles_m, cudaStream_t s) {
  using C = Cfg<BLOCK_N, RES>;
  static bool attr_set = false;
  if (!attr_set) {
    NSAC_CUDA(cudaFuncSetAttribute(gemm_bf16x3_kernel<BLOCK_N, NSAC_SPLIT_F16, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    NSAC_CUDA(cudaFuncSetAttribute(gemm_bf16x3_kernel<BLOCK_N, NSAC_SPLIT_BF16, RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int tiles = tiles_m * nsac_cdiv(p.N, BLOCK_N);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  if (p.fmt == NSAC_SPLIT_BF16)
    gemm_bf16x3_kernel<BLOCK_N, NSAC_SPLIT_BF16, RES><<<grid, C::THREADS, C::SMEM_BYTES, s>>>(ah, al, wh, wl, rh, rl, oh, ol, p);
  else
    gemm_bf16x3_kernel<BLOCK_N, NSAC_SPLIT_F16, RES><<<grid, C::THREADS, C::SMEM_BYTES, s>>>(ah, al, wh, wl, rh, rl, oh, ol, p);
  NSAC_CHECK_LAUNCH("nsac_gemm_split");
  return NSAC_OK;
}

// NHWC 16-bit planes [N,H,W,C] as a 4-D tensor (C, W, H, N); box = [64 ch, BW, BH, 1], 128-byte swizzle
bool make_map_nhwc(CUtensorMap* map, const void* base, int N, int H, int W, int C, int BW, int BH, bool is_bf16, int stride = 1) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  // traversal stride s: the box spans BW * s (BH * s) input pixels and TMA loads every s-th of them = BW (BH) elements
  cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(BW * stride), (cuuint32_t)(BH * stride), 1};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  return enc(map, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims,
             strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

// 3x3 / stride 1 / pad 1 convolution as an implicit GEMM: x planes [N,H,W,Cin] (Cin % 64 == 0), weight planes
// [Cout, 9*Cin] in (ky, kx, cin) order; output rows are NHWC pixels [N*H*W, Cout].
static unsigned long long* g_gemm_trace = nullptr;
// Profiling aid: while `buf` (16 uint64 device words) is set, CTA 0 of every GEMM-engine launch records %globaltimer at
// kernel entry, after setup, first TMA issue, first operands landed, first chunk committed, epilogue start / end, exit.
extern "C" int nsac_debug_gemm_trace(void* buf) {
  g_gemm_trace = static_cast<unsigned long long*>(buf);
  return NSAC_OK;
}

static int conv3x3_impl(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                        int N, int Hin, int Win, int Cin, int Cout, int stride, int act, int passes, int fmt, float out_scale,
                        float* out_f32, int ldo, void* out_hi, void* out_lo, int ld_split, void* stream, int taps = 9) {
  NSAC_REQUIRE(stride == 1 || stride == 2, "nsac_conv3x3_split: stride must be 1 or 2");
  const int H = (Hin - 1) / stride + 1, W = (Win - 1) / stride + 1;      // output map (3x3, pad 1)
  NSAC_REQUIRE(x_hi && w_hi, "nsac_conv3x3_split: null operand");
  NSAC_REQUIRE(passes >= 1 && passes <= 4 && (passes < 2 || x_lo) && (passes < 3 || w_lo), "nsac_conv3x3_split: bad passes / planes");
  NSAC_REQUIRE(N >= 0 && H >= 1 && W >= 1 && Cin >= 64 && Cin % 64 == 0 && Cout >= 8, "nsac_conv3x3_split: bad shape (Cin %% 64 == 0)");
  NSAC_REQUIRE(out_f32 || out_hi, "nsac_conv3x3_split: no output requested");
  NSAC_REQUIRE(!out_f32 || (ldo >= Cout && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0), "nsac_conv3x3_split: fp32 output alignment");
  NSAC_REQUIRE(!out_hi || (out_lo && ld_split >= Cout && ld_split % 8 == 0), "nsac_conv3x3_split: split output layout");
  NSAC_REQUIRE(fmt == NSAC_SPLIT_F16 || fmt == NSAC_SPLIT_BF16, "nsac_conv3x3_split: bad plane format");
  if (N == 0) return NSAC_OK;
  // pixel patch of one M tile: BW x BH <= 128 rows
  int BW, BH;
  if (W <= 128) { BW = W; BH = 128 / W; if (BH > H) BH = H; }
  else { BW = 128; BH = 1; }
  if (W % 16 == 0 && W > 64) { BW = 16; BH = 8; }          // e.g. 80-wide maps: 16 x 8 patches fill all 128 rows
  NSAC_REQUIRE(BW * stride <= 256 && BH * stride <= 256 && BW * BH <= 128, "nsac_conv3x3_split: cannot tile a %d x %d map", H, W);
  const bool bf = fmt == NSAC_SPLIT_BF16;
  const int K = taps * Cin;
  CUtensorMap mah, mal, mwh, mwl;
  const bool ok = make_map_nhwc(&mah, x_hi, N, Hin, Win, Cin, BW, BH, bf, stride) &&
                  make_map_nhwc(&mal, x_lo ? x_lo : x_hi, N, Hin, Win, Cin, BW, BH, bf, stride) &&
                  make_map(&mwh, w_hi, Cout, K, K, Cout <= 64 ? 64 : 128, bf) && make_map(&mwl, w_lo ? w_lo : w_hi, Cout, K, K, Cout <= 64 ? 64 : 128, bf);
  if (!ok) {
    nsac_set_error("nsac_conv3x3_split: cuTensorMapEncodeTiled failed (N=%d H=%d W=%d Cin=%d Cout=%d)", N, H, W, Cin, Cout);
    return NSAC_ERR_LAUNCH;
  }
  GemmParams p;
  p.trace = g_gemm_trace;
  p.bias = bias; p.bias_group_rows = 0; p.M = N * H * W; p.N = Cout; p.K = K; p.act = act; p.passes = passes;
  p.fmt = fmt; p.out_scale = out_scale; p.out_f32 = out_f32; p.ldo = ldo;
  p.out_hi = static_cast<uint16_t*>(out_hi); p.out_lo = static_cast<uint16_t*>(out_lo); p.ld_split = ld_split;
  p.res_hi = nullptr; p.res_lo = nullptr; p.ld_res = 0; p.a_lo_zero = 0; p.row_bias = nullptr;
  p.conv = 1; p.conv_taps = taps; p.H = H; p.W = W; p.BW = BW; p.BH = BH; p.cblocks = Cin / 64; p.cstride = stride;
  p.tiles_x = nsac_cdiv(W, BW); p.tiles_y = nsac_cdiv(H, BH);
  if (Cout <= 64) return launch_gemm<64, 0>(mah, mal, mwh, mwl, mwh, mwl, mwh, mwl, p, N * p.tiles_x * p.tiles_y, static_cast<cudaStream_t>(stream));
  return launch_gemm<128, 0>(mah, mal, mwh, mwl, mwh, mwl, mwh, mwl, p, N * p.tiles_x * p.tiles_y, static_cast<cudaStream_t>(stream));
}

extern "C" int nsac_conv3x3_split(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                                  int N, int H, int W, int Cin, int Cout, int act, int passes, int fmt, float out_scale,
                                  float* out_f32, int ldo, void* out_hi, void* out_lo, int ld_split, void* stream) {
  return conv3x3_impl(x_hi, x_lo, w_hi, w_lo, bias, N, H, W, Cin, Cout, 1, act, passes, fmt, out_scale, out_f32, ldo, out_hi, out_lo,
                      ld_split, stream);
}

// Same with a convolution stride of 1 or 2 (H, W = INPUT map; output (H-1)/stride+1 x (W-1)/stride+1): the strided 3x3 of
// res3.0 / res4.0 / res5.0 (STRIDE_IN_1X1 = False) as an implicit GEMM - the A tiles are gathered by TMA with a traversal
// stride, no im2col matrix.
extern "C" int nsac_conv3x3_split_strided(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                                          int N, int H, int W, int Cin, int Cout, int stride, int act, int passes, int fmt,
                                          float out_scale, float* out_f32, int ldo, void* out_hi, void* out_lo, int ld_split,
                                          void* stream) {
  return conv3x3_impl(x_hi, x_lo, w_hi, w_lo, bias, N, H, W, Cin, Cout, stride, act, passes, fmt, out_scale, out_f32, ldo, out_hi,
                      out_lo, ld_split, stream);
}

// 1x1 convolution with a stride (no padding): the projection shortcut of res3.0 / res4.0 / res5.0 reads every second pixel of
// every second row straight through the tensor map's traversal stride - no subsampled copy of the input.
extern "C" int nsac_conv1x1_split_strided(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, const float* bias,
                                          int N, int H, int W, int Cin, int Cout, int stride, int act, int passes, int fmt,
                                          float out_scale, float* out_f32, int ldo, void* out_hi, void* out_lo, int ld_split,
                                          void* stream) {
  return conv3x3_impl(x_hi, x_lo, w_hi, w_lo, bias, N, H, W, Cin, Cout, stride, act, passes, fmt, out_scale, out_f32, ldo, out_hi,
                      out_lo, ld_split, stream, 1);
}

static int gemm_split_impl(const void* a_hi, const void* a_lo, int lda, const void* w_hi, const void* w_lo,
                           int ldw, const float* bias, int bias_group_rows, int M, int N, int K, int act,
                           int passes, int fmt, float out_scale, float* out_f32, int ldo, void* out_hi,
                           void* out_lo, int ld_split, const void* res_hi, const void* res_lo, int ld_res, const float* row_bias,
                           void* stream) {
  NSAC_REQUIRE(!res_hi || (res_lo && ld_res >= N && ld_res % 8 == 0 && (reinterpret_cast<uintptr_t>(res_hi) & 15) == 0 &&
                           (reinterpret_cast<uintptr_t>(res_lo) & 15) == 0),
               "nsac_gemm_split_residual: residual needs both planes, 16-byte alignment and ld_res %% 8 == 0");
  NSAC_REQUIRE(!res_hi || (out_hi && !out_f32 && bias_group_rows == 0),
               "nsac_gemm_split_residual: the residual epilogue writes hi/lo planes only (out_hi/out_lo, no out_f32, no grouped bias)");
  NSAC_REQUIRE(a_hi && w_hi, "nsac_gemm_split: null operand");
  NSAC_REQUIRE(passes >= 1 && passes <= 4, "nsac_gemm_split: passes must be 1..4");
  NSAC_REQUIRE(passes < 3 || w_lo, "nsac_gemm_split: missing W lo plane for %d passes", passes);     // a_lo == nullptr: A has no lo plane
  NSAC_REQUIRE(M >= 0 && N >= 8 && K >= BLOCK_K && K % BLOCK_K == 0, "nsac_gemm_split: need K %% 64 == 0 (M=%d N=%d K=%d)", M, N, K);
  NSAC_REQUIRE(lda >= K && ldw >= K && lda % 8 == 0 && ldw % 8 == 0, "nsac_gemm_split: lda/ldw must be >= K and multiples of 8");
  NSAC_REQUIRE(out_f32 || out_hi, "nsac_gemm_split: no output requested");
  NSAC_REQUIRE(!out_f32 || (ldo >= N && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0),
               "nsac_gemm_split: fp32 output must be 16-byte aligned with ldo %% 4 == 0");
  NSAC_REQUIRE(!out_hi || (out_lo && ld_split >= N && ld_split % 8 == 0 && (reinterpret_cast<uintptr_t>(out_hi) & 15) == 0 &&
                           (reinterpret_cast<uintptr_t>(out_lo) & 15) == 0),
               "nsac_gemm_split: split output needs both planes, 16-byte alignment and ld_split %% 8 == 0");
  NSAC_REQUIRE(act >= 0 && act <= 2, "nsac_gemm_split: bad act %d", act);
  NSAC_REQUIRE(fmt == NSAC_SPLIT_F16 || fmt == NSAC_SPLIT_BF16, "nsac_gemm_split: bad plane format %d", fmt);
  for (const void* ptr : {a_hi, a_lo, w_hi, w_lo})
    NSAC_REQUIRE(!ptr || (reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "nsac_gemm_split: operands must be 16-byte aligned");
  if (M == 0) return NSAC_OK;
  const int block_n = N <= 64 ? 64 : 128;      // 64-wide tiles for the narrow layers (res2 / stem of the backbone): no half-empty MMAs
  CUtensorMap mah, mal, mwh, mwl;
  const bool bf = fmt == NSAC_SPLIT_BF16;
  bool ok = make_map(&mah, a_hi, M, K, lda, BLOCK_M, bf) && make_map(&mal, a_lo ? a_lo : a_hi, M, K, lda, BLOCK_M, bf) &&
            make_map(&mwh, w_hi, N, K, ldw, block_n, bf) && make_map(&mwl, w_lo ? w_lo : w_hi, N, K, ldw, block_n, bf);
  if (!ok) {
    nsac_set_error("nsac_gemm_split: cuTensorMapEncodeTiled failed (M=%d N=%d K=%d lda=%d ldw=%d)", M, N, K, lda, ldw);
    return NSAC_ERR_LAUNCH;
  }
  GemmParams p;
  p.trace = g_gemm_trace;
  p.bias = bias; p.bias_group_rows = bias_group_rows; p.M = M; p.N = N; p.K = K; p.act = act; p.passes = passes;
  p.fmt = fmt; p.out_scale = out_scale;
  p.out_f32 = out_f32; p.ldo = ldo;
  p.out_hi = static_cast<uint16_t*>(out_hi); p.out_lo = static_cast<uint16_t*>(out_lo); p.ld_split = ld_split;
  p.res_hi = static_cast<const uint16_t*>(res_hi); p.res_lo = static_cast<const uint16_t*>(res_lo); p.ld_res = ld_res;
  p.a_lo_zero = a_lo == nullptr ? 1 : 0;
  p.row_bias = row_bias;
  p.conv = 0; p.conv_taps = 1; p.H = p.W = p.BW = p.BH = p.cblocks = p.tiles_x = p.tiles_y = 1; p.cstride = 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (res_hi) {
    // residual epilogue: 64-wide tiles, residual tiles TMA-prefetched and turned in place into the output tile, TMA store
    CUtensorMap mrh, mrl, moh, mol, mw64h, mw64l;
    if (!(make_map(&mrh, res_hi, M, N, ld_res, BLOCK_M, bf) && make_map(&mrl, res_lo, M, N, ld_res, BLOCK_M, bf) &&
          make_map(&moh, out_hi, M, N, ld_split, BLOCK_M, bf) && make_map(&mol, out_lo, M, N, ld_split, BLOCK_M, bf) &&
          make_map(&mw64h, w_hi, N, K, ldw, 64, bf) && make_map(&mw64l, w_lo ? w_lo : w_hi, N, K, ldw, 64, bf))) {
      nsac_set_error("nsac_gemm_split_residual: cuTensorMapEncodeTiled failed (M=%d N=%d ld_res=%d ld_split=%d)", M, N, ld_res, ld_split);
      return NSAC_ERR_LAUNCH;
    }
    // K <= 128 (res2 / res3 of the backbone: one or two k-blocks per tile, HBM-bound): the smem budget goes to FOUR tile buffers
    // (residual prefetch + store drain depth: 0.69 -> 0.88 of the HBM bound) and two operand stages; deeper K keeps three
    // operand stages and two tile buffers (operand delivery is what bounds those; measured both ways, profiles/README.md r2y)
    if (K <= 128) return launch_gemm<64, 2>(mah, mal, mw64h, mw64l, mrh, mrl, moh, mol, p, nsac_cdiv(M, BLOCK_M), s);
    // K >= 512 (res5): 128-wide tiles halve the A traffic from L2 and the MMAs of a tile last long enough to hide the heavier
    // epilogue (residual fetched from global through the staging buffer): 341 -> 295 us; at K = 256 (res4) that epilogue is
    // the limiter (482 vs 424 us) and the 64-wide in-place variant stays (profiles/README.md, r2dd / r2ee)
    if (block_n == 128 && K >= 512)
      return launch_gemm<128, 3>(mah, mal, mwh, mwl, mwh, mwl, mwh, mwl, p, nsac_cdiv(M, BLOCK_M), s);
    return launch_gemm<64, 1>(mah, mal, mw64h, mw64l, mrh, mrl, moh, mol, p, nsac_cdiv(M, BLOCK_M), s);
  }
  if (block_n == 64) return launch_gemm<64, 0>(mah, mal, mwh, mwl, mwh, mwl, mwh, mwl, p, nsac_cdiv(M, BLOCK_M), s);
  return launch_gemm<128, 0>(mah, mal, mwh, mwl, mwh, mwl, mwh, mwl, p, nsac_cdiv(M, BLOCK_M), s);
}

extern "C" int nsac_gemm_split(const void* a_hi, const void* a_lo, int lda, const void* w_hi, const void* w_lo,
                               int ldw, const float* bias, int bias_group_rows, int M, int N, int K, int act,
                               int passes, int fmt, float out_scale, float* out_f32, int ldo, void* out_hi,
                               void* out_lo, int ld_split, void* stream) {
  return gemm_split_impl(a_hi, a_lo, lda, w_hi, w_lo, ldw, bias, bias_group_rows, M, N, K, act, passes, fmt, out_scale, out_f32, ldo,
                         out_hi, out_lo, ld_split, nullptr, nullptr, 0, nullptr, stream);
}

// nsac_gemm_split + a per-ROW scalar: out[m, n] = act(out_scale * acc + bias[n] + row_bias[m]).
extern "C" int nsac_gemm_split_rowbias(const void* a_hi, const void* a_lo, int lda, const void* w_hi, const void* w_lo,
                                       int ldw, const float* bias, const float* row_bias, int M, int N, int K, int act,
                                       int passes, int fmt, float out_scale, float* out_f32, int ldo, void* out_hi,
                                       void* out_lo, int ld_split, void* stream) {
  return gemm_split_impl(a_hi, a_lo, lda, w_hi, w_lo, ldw, bias, 0, M, N, K, act, passes, fmt, out_scale, out_f32, ldo, out_hi, out_lo,
                         ld_split, nullptr, nullptr, 0, row_bias, stream);
}

// nsac_gemm_split + a residual input given as hi/lo planes [M, ld_res]: out = act(A.W^T * scale + bias + residual) - the
// `out += shortcut; relu` of a bottleneck block (detectron2 BottleneckBlock.forward) inside the producing GEMM's epilogue.
extern "C" int nsac_gemm_split_residual(const void* a_hi, const void* a_lo, int lda, const void* w_hi, const void* w_lo,
                                        int ldw, const float* bias, int M, int N, int K, int act, int passes, int fmt,
                                        float out_scale, const void* res_hi, const void* res_lo, int ld_res, float* out_f32,
                                        int ldo, void* out_hi, void* out_lo, int ld_split, void* stream) {
  NSAC_REQUIRE(res_hi && res_lo, "nsac_gemm_split_residual: null residual planes");
  return gemm_split_impl(a_hi, a_lo, lda, w_hi, w_lo, ldw, bias, 0, M, N, K, act, passes, fmt, out_scale, out_f32, ldo, out_hi, out_lo,
                         ld_split, res_hi, res_lo, ld_res, nullptr, stream);
}

extern "C" int nsac_split16(const float* x, int ldx, int rows, int K, float scale, int fmt, void* hi, void* lo,
                            int ld_split, void* stream) {
  NSAC_REQUIRE(x && hi && lo, "nsac_split16: null pointer");
  NSAC_REQUIRE(rows >= 0 && K >= 1 && ldx >= K && ld_split >= K, "nsac_split16: bad shape");
  NSAC_REQUIRE(fmt == NSAC_SPLIT_F16 || fmt == NSAC_SPLIT_BF16, "nsac_split16: bad plane format %d", fmt);
  if (rows == 0) return NSAC_OK;
  const size_t total = (size_t)rows * ld_split;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  split16_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ldx, rows, K, scale, fmt, static_cast<uint16_t*>(hi), static_cast<uint16_t*>(lo), ld_split);
  NSAC_CHECK_LAUNCH("nsac_split16");
  return NSAC_OK;
}

unsigned int* nsac_pixel_overflow_ptr();      // csrc/pixel.cu

// Sticky fp16-plane overflow flag: *flag_out = 1 if, since the last clearing call, nsac_split16 / nsac_nchw_to_planes / a GEMM
// epilogue had to write a finite value with |x| > 65504 into an fp16 plane (it became inf there: every result computed from it
// is invalid).  Synchronises `stream`; `clear` != 0 resets the flag.  Debug / validation aid, not on the hot path.
extern "C" int nsac_plane_overflow(int* flag_out, int clear, void* stream) {
  NSAC_REQUIRE(flag_out, "nsac_plane_overflow: null pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned int* ptrs[2] = {nullptr, nsac_pixel_overflow_ptr()};
  NSAC_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&ptrs[0]), g_plane_overflow));
  unsigned int v[2] = {0u, 0u};
  for (int i = 0; i < 2; ++i) {
    if (!ptrs[i]) continue;
    NSAC_CUDA(cudaMemcpyAsync(&v[i], ptrs[i], sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    if (clear) NSAC_CUDA(cudaMemsetAsync(ptrs[i], 0, sizeof(unsigned int), s));
  }
  NSAC_CUDA(cudaStreamSynchronize(s));
  *flag_out = (v[0] | v[1]) ? 1 : 0;
  return NSAC_OK;
}
