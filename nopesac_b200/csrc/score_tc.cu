// K8 + K9 on the Blackwell tensor pipe — hypothesis scoring and pose selection (camera_head.py:964-1115).
//
// Three kernels per call:
//   score_prep_kernel        per pair: column blocks (residual constants + HMMA B fragments) of every k-block of 64
//                            matched columns, HMMA A fragments of every hypothesis row, hypothesis 0 (the initial pose)
//   score_tc_kernel          persistent, one CTA per SM, 32 warps in 8 warpgroups (setmaxnreg), work item = (pair b, tile
//                            of 128 one-plane hypotheses h = 1 + 128*tile + r) or a row-0 tile (hypothesis 0 of 128 pairs):
//     warps 8-23   residuals  u = (kR) n^ and t.u on HMMA (fp16 hi/lo k-slots), then per row x column pair 13 packed fp32
//                             instructions + 4 MUFU.SQRT + 4 MUFU.EX2:  rot: exp(-|u - n1|), trans: exp(-|A (d + t.u) u - pi1|)
//                             (closed forms of the reference's warp + normalise), written as fp16 straight into the
//                             128-byte-swizzled K-major A-operand tiles (2-stage ring), fence.proxy.async, mbarrier arrive.
//                             The [B,NQ+1,NQ,3] temporaries of the reference never exist.
//     warps 0-7    gather     flash-style softmax partials: local max / sum of exp over the tile's logits and the
//                             exp-weighted + plain sums of the tile's [128,256] one-plane features, read from the feature ring
//     warps 24-27  epilogue   TMEM -> registers: +b1, ReLU, fp16x2 -> tcgen05.st (H1, in place);  then +b2, ReLU and the
//                             folded Linear(128,64)+Linear(64,1) dot product: one thread owns one hypothesis row
//     warp 28      TMA        W1 k-blocks (rot / trans [128 x 64] fp16 slices), then the tile's two W2 k-blocks (2-stage ring)
//     warp 29      MMA        layer 1: D_b[128x128] += X_b[128x64] . W1_b^T (tcgen05.mma kind::f16, A and B from shared
//                             memory);  layer 2: D2_b = H1_b . W2_b^T with H1 read from TENSOR MEMORY
//     warp 30      TMA        column blocks (CJ ring) + row-0 tiles' A stages
//     warp 31      TMA        feature ring: the only HBM stream of the kernel
//   score_select_tc_kernel   per pair: merges the tile partials with hypothesis 0 (log-sum-exp rescale), applies the
//                            m == 0 / m == 1 / m > 1 rules (:964, :1052, :1068), the avg / soft / min-cost / max-score
//                            selection and the shared pose heads, and writes pose[b, 0:16].
//
// Precision: the score MLPs run single-pass fp16 (11 significant bits for exp(-d) in [0,1] and for the
// weights) with fp32 accumulation; measured effect on scores <= 4e-5 and on the soft pose <= 3e-5 even with a
// 4x sharpened softmax.  The exact-fp32 CUDA-core kernels of score.cu remain available
// (precision = 0, and always for the diagnostic outputs).
#include <cuda.h>
#include <cuda_fp16.h>
#include <string.h>
#include "common.cuh"

#ifndef NSAC_SCORE_L2PREFETCH
#define NSAC_SCORE_L2PREFETCH 0
#endif
#ifndef NSAC_SCORE_ABLATE       // profiling builds only (scripts/build_variant.sh): 1 = MUFU ops replaced by multiplies, 2 = HMMAs skipped,
#define NSAC_SCORE_ABLATE 0     // 3 = residual warps do no math at all, 4 = gather warps skip the feature math, 5 = tcgen05 MMAs skipped.
#endif                          // Results are WRONG in these builds; they only time what is left.
#ifndef NSAC_SCORE_PDL          // programmatic dependent launch between prep -> tiles -> selection
#define NSAC_SCORE_PDL 1
#endif
#ifndef NSAC_SCORE_SWPIPE       // residual groups software-pipelined by hand (next group's HMMAs before this group's MUFUs)
#define NSAC_SCORE_SWPIPE 0      // r2m A/B on one box: 104.2 us without vs 106.5 us with - the kernel is shared-memory-bandwidth bound, not latency bound
#endif

namespace {

constexpr int TILE_H = 128;        // hypotheses per tile (UMMA M)
constexpr int HID = 128;           // hidden width of the score MLPs (UMMA N)
constexpr int KB = 64;             // residual columns per k-block (one 128-byte swizzle row of fp16)
constexpr int C_FEAT = 256;
#ifndef NSAC_SCORE_W_STAGES
#define NSAC_SCORE_W_STAGES 2
#endif
#ifndef NSAC_SCORE_F_STAGES
#define NSAC_SCORE_F_STAGES 4
#endif
#ifndef NSAC_SCORE_A_STAGES
#define NSAC_SCORE_A_STAGES 2
#endif
constexpr int W_STAGES = NSAC_SCORE_W_STAGES, A_STAGES = NSAC_SCORE_A_STAGES;
constexpr int BLK_BYTES = TILE_H * KB * 2;           // 16 KB: one [128 x 64] fp16 operand block
// warp roles, aligned to warpgroups of 4 warps so that setmaxnreg can move registers between the roles
// (launch: 1024 threads x 64 registers = the whole register file):
//   WG0-1 = warps 0-7   gather              40 regs   | WG2-5 = warps 8-23   residuals (16)      80 regs
//   WG6   = warps 24-27 epilogue            72 regs   | WG7   = warps 28-31  W / MMA / column / feature producers  40 regs
// setmaxnreg.inc can only draw on what the CTA's own warps released with setmaxnreg.dec (an inc that is not covered
// deadlocks): released 128*24 + 256*24 = 9216 = claimed 128*8 + 512*16.
// Four residual warps per scheduler: the FMA-pipe phase (packed FFMA2-class instructions, 2 issue cycles each) of one
// warp overlaps the MUFU phase (8 SQRT/EX2 per row x column pair, 8 cycles each) of another - with two warps per
// scheduler (r1e/s1 captures) the two pipes alternated instead of overlapping and neither was more than 38 % busy.
// The warp scheduler prefers the HIGHEST warp id among eligible warps (B300_MICROARCH.md "Multi-warp arbiter"), so the
// latency-critical single-thread roles sit at the top, then the epilogue (on the MMA issuer's critical chain), then the
// residual warps (the kernel's critical path), and the gather warps - which have slack - at the bottom.
constexpr int NUM_WARPS = 32, NUM_THREADS = NUM_WARPS * 32;
constexpr int G_WARP0 = 0, G_THREADS = 256;
constexpr int R_WARP0 = 8, R_WARPS = 16, R_THREADS = R_WARPS * 32;
constexpr int E_WARP0 = 24;
constexpr int T_WARP = 28, M_WARP = 29, C_WARP = 30, F_WARP = 31;
constexpr uint32_t SPIN_LIMIT = 2000;      // suspended waits of up to ~10 ms each: ~20 s, then trap instead of hanging
constexpr int PART_HDR = 4;                          // max, sumexp, pad, pad (keeps the vectors 16-byte aligned)
constexpr int PART_STRIDE = PART_HDR + 2 * C_FEAT;   // per (pair, tile, branch): header, wsum[256], fsum[256]

// shared memory map (offsets from a 1024-aligned base)
constexpr int F_STAGES = NSAC_SCORE_F_STAGES, F_BYTES = 16384;                       // feature ring: 16 hypothesis rows x 256 fp32 per stage (the single
                                                                   // producer thread needs ~0.26 us per bulk copy: 8 KB chunks capped the stream at 31 GB/s per SM)
constexpr int F_ROWS = F_BYTES / (C_FEAT * 4);
constexpr int OFF_F = 0;                                         // [F_STAGES][F_BYTES] feature ring (TMA bulk copies)
constexpr int OFF_W1 = OFF_F + F_STAGES * F_BYTES;               // [stage][branch][16 KB]: W1 k-blocks, then the tile's two W2 k-blocks
constexpr int OFF_A = OFF_W1 + W_STAGES * 2 * BLK_BYTES;         // [stage][branch][16 KB]
constexpr int CJK_BYTES = 4096;                                   // per k-block: 2 KB column constants + 2 KB B fragments
constexpr int OFF_CJ = OFF_A + A_STAGES * 2 * BLK_BYTES;         // [A_STAGES][CJK_BYTES] column block of the k-block (TMA)
constexpr int OFF_VEC = OFF_CJ + A_STAGES * CJK_BYTES;           // b1[2][128], b2[2][128], w34[2][128] floats
constexpr int OFF_LOGIT = OFF_VEC + 6 * HID * 4;                 // [2 bufs][2 branches][128] floats
constexpr int OFF_ROWSUM = OFF_LOGIT + 2 * 2 * TILE_H * 4;       // [2 column halves][2 branches][128] floats (min-cost sums)
constexpr int OFF_EXP = OFF_ROWSUM + 2 * 2 * TILE_H * 4;         // [2 branches][128] softmax numerators of the tile
constexpr int OFF_GPART = OFF_EXP + 2 * TILE_H * 4;              // [2 branches][64 threads][9] odd-row partials (+1 pad)
constexpr int OFF_BAR = OFF_GPART + 2 * 64 * 12 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
static_assert(F_STAGES % 2 == 0, "rot / tran chunks alternate through the feature ring");

// tensor-memory columns
constexpr uint32_t TM_D = 0;          // layer 1: D_rot [0,128), D_tran [128,256); H1 (fp16 pairs) overwrites the first 64
                                      // columns of each branch IN PLACE (a thread only touches its own lane, after reading it)
constexpr uint32_t TM_D2 = 256;       // layer 2: D2_rot [256,384), D2_tran [384,512) - so the next tile's layer-1 MMAs run
                                      // while the epilogue warps still read D2 (MMAs execute in issue order: no barrier)
constexpr uint32_t TM_COLS = 512;

enum { BAR_CJ_FULL = 0, BAR_CJ_EMPTY = BAR_CJ_FULL + A_STAGES, BAR_W_FULL = BAR_CJ_EMPTY + A_STAGES, BAR_W_EMPTY = BAR_W_FULL + W_STAGES,
       BAR_A_FULL = BAR_W_EMPTY + W_STAGES, BAR_A_EMPTY = BAR_A_FULL + A_STAGES, BAR_ACC_FULL = BAR_A_EMPTY + A_STAGES,
       BAR_H1_READY = BAR_ACC_FULL + 1, BAR_ACC2_FULL = BAR_H1_READY + 1, BAR_LOGIT_READY = BAR_ACC2_FULL + 1,
       BAR_LOGIT_FREE = BAR_LOGIT_READY + 2, BAR_F_FULL = BAR_LOGIT_FREE + 2, BAR_F_EMPTY = BAR_F_FULL + F_STAGES,
       BAR_COUNT = BAR_F_EMPTY + F_STAGES };
static_assert(BAR_COUNT * 8 + 8 <= 512, "barrier area");

// ------------------------------------------------------------------------------------------------ PTX helpers
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
        : "=r"(done) : "r"(addr), "r"(parity), "r"(0x989680u) : "memory");
    if (done) break;
    if (++spins > SPIN_LIMIT) __trap();
  }
}
// Fine-grained rings (8 KB feature chunks, ~0.2 us apart): plain try_wait polling, no suspend hint - a parked warp
// resumes too late for this cadence.  Only the 8 gather warps and the feature producer use it.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) break;
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// L2 prefetch of a contiguous global range (no shared-memory destination, no barrier): decouples the HBM stream from the
// shared-memory ring, which can only run 64 KB ahead of its consumers
__device__ __forceinline__ void l2_prefetch(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(src)), "r"(bytes) : "memory");
}
// programmatic dependent launch: let the next kernel of the stream start its prologue / wait for the previous one's results
#if NSAC_SCORE_PDL
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#else
__device__ __forceinline__ void pdl_launch_dependents() {}
__device__ __forceinline__ void pdl_wait() {}
#endif
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// K-major SWIZZLE_128B shared-memory descriptor (see gemm_tc.cu)
__device__ __forceinline__ uint64_t sw128_desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// fp16 x fp16 -> fp32, K-major A and B, M = 128, N = 128
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)(TILE_H >> 4) << 24);

// ------------------------------------------------------------------------------------------------ residual math
// Reference (camera_head.py:997-1035 with the warp of :1446-1453): e = R (p0*flip) + t, b = e - t, pi0 = (e.b/(|b|+1e-5)^2) b.
// With u = R n^ (unit) and d = |p0|: b = d u, e.b = d^2 + d (t.u), so pi0 = A (d + t.u) u; for t = 0 its direction is u.
//   rot  : exp(-| u - n1 |)                 (F.normalize of both sides)
//   trans: exp(-| A (d + t.u) u - pi1 |)
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_exp2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// ---- packed fp32 (sm_100 FFMA2 / FMUL2 / FADD2: two fp32 lanes per issued instruction).  The residual warps are
// issue-bound (ncu r1e: 61.6 M warp instructions, 55 % issue-slot utilisation), so the math is laid out on column
// PAIRS (j, j+1): the per-column constants arrive interleaved, the per-row constants (R, t) use the scalar-
// broadcast operand form, and the two results convert to one fp16x2 word with a single cvt.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 bc2(float s) { return pk2(s, s); }       // ptxas folds this into the .F32 broadcast operand
// fp16x2 word {lo -> bits 0-15, hi -> bits 16-31}; RELU clamps negatives to 0 inside the conversion
template <bool RELU>
__device__ __forceinline__ uint32_t cvt_h2(float lo, float hi) {
  uint32_t r;
  if (RELU) asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  else      asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

constexpr float LOG2E = 1.4426950408889634f;
constexpr float CJ_BIG = 1000.f;      // |.| of a padded column in log2 units: ex2(-1000) == 0 (ftz)
constexpr int CJ_FIELDS = 13;         // per column: n^ (0-2), -k n1 (3-5), -k pi1 (6-8), A (9), Bc (10), valid (11), |pi1| (12); k = log2(e)

// Column constants of matched plane pair j from the geo_local row (p0, p1); layout in memory is per column PAIR:
// [pair][field][2].  Everything the exponent needs is pre-scaled by k = log2(e) so that x = ex2(-|.|) directly:
//   rot  : |k u - k n1|                  with k u = (k R) n^
//   trans: |g (k u) - k pi1|,  g = A (t.u) + A d,  t.u = (t / k).(k u)
// Padded columns (j >= m) get n^ = 0 and a BIG offset, so both results are exactly 0 without a mask multiply.
__device__ __forceinline__ void column_fields(const float* __restrict__ g6, bool valid, float (&f)[CJ_FIELDS]) {
#pragma unroll
  for (int i = 0; i < CJ_FIELDS; ++i) f[i] = 0.f;
  if (valid) {
    float ax = g6[0], ay = -g6[1], az = -g6[2];
    const float d = normalize3(ax, ay, az);
    const float px = g6[3], py = -g6[4], pz = -g6[5];
    float nx = px, ny = py, nz = pz;
    const float d1 = normalize3(nx, ny, nz);
    const float dd = d + 1e-5f, A = d * d / (dd * dd);
    f[0] = ax; f[1] = ay; f[2] = az;
    f[3] = -LOG2E * nx; f[4] = -LOG2E * ny; f[5] = -LOG2E * nz;
    f[6] = -LOG2E * px; f[7] = -LOG2E * py; f[8] = -LOG2E * pz;
    f[9] = A; f[10] = A * d; f[11] = 1.f; f[12] = d1;
  } else {
    f[3] = CJ_BIG; f[6] = CJ_BIG; f[12] = 1.f;       // (|pi1| = 1 keeps d1 * (-k n1) = BIG for the padded columns)
  }
}
// One hypothesis row against one column pair.  R = k * rotation, t = translation / k (see above).
// Reference (camera_head.py:997-1035 with the warp of :1446-1453): e = R (p0*flip) + t, b = e - t, pi0 = (e.b/(|b|+1e-5)^2) b.
// With u = R n^ (unit) and d = |p0|: b = d u, e.b = d^2 + d (t.u), so pi0 = A (d + t.u) u; for t = 0 its direction is u.
//   rot  : exp(-| u - n1 |)                 (F.normalize of both sides)
//   trans: exp(-| A (d + t.u) u - pi1 |)
// scalar form (score_prep_kernel: hypothesis 0 of every pair), one column
template <bool SUMS>
__device__ __forceinline__ void residual_scalar(const float (&R)[9], const float (&t)[3], const float (&c)[CJ_FIELDS], float& xr,
                                                float& xt, float& sum_r, float& sum_t) {
  const float ux = fmaf(R[0], c[0], fmaf(R[1], c[1], R[2] * c[2]));
  const float uy = fmaf(R[3], c[0], fmaf(R[4], c[1], R[5] * c[2]));
  const float uz = fmaf(R[6], c[0], fmaf(R[7], c[1], R[8] * c[2]));
  const float ax = ux + c[3], ay = uy + c[4], az = uz + c[5];
  const float dr = fast_sqrt(fmaf(ax, ax, fmaf(ay, ay, az * az)));
  const float tu = fmaf(t[0], ux, fmaf(t[1], uy, t[2] * uz));
  const float g = fmaf(tu, c[9], c[10]);
  const float wx = fmaf(g, ux, c[6]), wy = fmaf(g, uy, c[7]), wz = fmaf(g, uz, c[8]);
  const float dt = fast_sqrt(fmaf(wx, wx, fmaf(wy, wy, wz * wz)));
  xr = fast_exp2(-dr);
  xt = fast_exp2(-dt);
  if (SUMS) { sum_r = fmaf(c[11], dr, sum_r); sum_t = fmaf(c[11], dt, sum_t); }
}

// scaled per-row constants: R <- k R, t <- t / k
__device__ __forceinline__ void row_constants(float qw, float qx, float qy, float qz, float tx, float ty, float tz, float (&R)[9],
                                              float (&t)[3]) {
  const Mat3 M = quat_to_rot(qw, qx, qy, qz);
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = LOG2E * M.m[i];
  t[0] = tx * (1.f / LOG2E); t[1] = ty * (1.f / LOG2E); t[2] = tz * (1.f / LOG2E);
}

// ---- u = (k R) n^ and t.u on the (legacy, warp-level) tensor path ------------------------------------------------------
// The 3x3 rotation of every (row, column) pair is 9 of the 25 packed FMA-pipe instructions of residual_cp, and the FMA
// pipe, the MUFU pipe and the issue slots are all ~equally loaded (s6 capture + scripts/ubench): the kernel is bound by
// their sum.  mma.sync.m16n8k16 (HMMA.16816.F32: 8 cycles per warp instruction per scheduler, overlaps MUFU completely,
// scripts/ubench/hmma.cu) computes u_c[16 rows x 8 columns] = A_c . B in one instruction with fp32-grade accuracy from
// fp16 hi/lo splits laid out along K:   k-slots  q=0: (hi0 hi1 | hi2 0)  q=1: (lo0 lo1 | lo2 0)  q=2: (hi0 hi1 | hi2 0)
//                                     B k-slots  q=0: (hi0 hi1 | hi2 0)  q=1: (hi0 hi1 | hi2 0)  q=2: (lo0 lo1 | lo2 0)
// i.e. hi.hi + lo.hi + hi.lo (the dropped lo.lo term is < 4e-7 absolute).  Thread (g = lane / 4, q = lane % 4) owns the
// k-slots (2q, 2q+1 | 2q+8, 2q+9) of rows g, g+8 (A) / column g (B) and receives rows g, g+8 x columns 2q, 2q+1 (D):
// exactly one packed column pair per row.
__device__ __forceinline__ void hmma16816(float (&d)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ uint32_t h_bits(float x) { return (uint32_t)__half_as_ushort(__float2half_rn(x)); }
// this thread's two A words (k-slots 2q,2q+1 and 2q+8,2q+9) of one row for the 3-vector v
__device__ __forceinline__ void a_words(float v0, float v1, float v2, int q, uint32_t& w01, uint32_t& w2) {
  if (q == 1) {      // lo parts
    v0 -= __half2float(__float2half_rn(v0)); v1 -= __half2float(__float2half_rn(v1)); v2 -= __half2float(__float2half_rn(v2));
  }
  w01 = h_bits(v0) | (h_bits(v1) << 16);
  w2 = h_bits(v2);
  if (q == 3) { w01 = 0u; w2 = 0u; }
}
constexpr int CJ8_FIELDS = 6;         // per column: -k n1 (0-2), |pi1| (3), A (4), Bc (5); -k pi1 = |pi1| * (-k n1) is rebuilt with 3
                                      // packed multiplies per group instead of being loaded: every one of these broadcast loads
                                      // costs 4 shared-memory wavefronts and the kernel is bound by that pipe (16 -> 12 per group).
                                      // Shared-memory layout: [group of 4 column
                                      // pairs][16-byte unit i = fields 2i, 2i+1][pair q][field parity][column parity] - the four pairs a
                                      // warp reads with one LDS.128 are 64 contiguous bytes (one wavefront; the [pair][field] layout
                                      // cost 8 wavefronts per load and the shared-memory pipe became the bottleneck, s7 timeline)

// everything after u, first half (FMA pipe): squared distances of one row x one column pair
__device__ __forceinline__ void residual_dist2(u64 ux, u64 uy, u64 uz, u64 tu, const u64 (&c)[CJ8_FIELDS + 3], u64& dr2, u64& dt2) {
  const u64 ax = fadd2(ux, c[0]), ay = fadd2(uy, c[1]), az = fadd2(uz, c[2]);
  dr2 = ffma2(ax, ax, ffma2(ay, ay, fmul2(az, az)));
  const u64 g = ffma2(tu, c[4], c[5]);
  const u64 wx = ffma2(g, ux, c[6]), wy = ffma2(g, uy, c[7]), wz = ffma2(g, uz, c[8]);      // c[6..8] = -k pi1 (rebuilt by the caller)
  dt2 = ffma2(wx, wx, ffma2(wy, wy, fmul2(wz, wz)));
}
// second half (MUFU pipe): 4 SQRT + 4 EX2 + 2 cvt.  The MUFU instructions are `asm volatile` like the HMMAs, so their order
// relative to the NEXT group's HMMAs is the source order (see residual_kblock_mma).
#if NSAC_SCORE_ABLATE == 1
__device__ __forceinline__ float vsqrt(float x) { return x * 0.5f; }
__device__ __forceinline__ float vexp2n(float x) { return x * 0.25f; }
#else
__device__ __forceinline__ float vsqrt(float x) { float r; asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float vexp2n(float x) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-x)); return r; }
#endif
template <bool SUMS>
__device__ __forceinline__ void residual_finish(u64 dr2, u64 dt2, u64 valid, uint32_t& hr, uint32_t& ht, u64& sum_r, u64& sum_t) {
  float r0, r1, t0, t1;
  upk2(dr2, r0, r1);
  upk2(dt2, t0, t1);
  r0 = vsqrt(r0); r1 = vsqrt(r1); t0 = vsqrt(t0); t1 = vsqrt(t1);
  if (SUMS) {   // masked distance sums (in log2 units; rescaled by the caller)
    sum_r = ffma2(valid, pk2(r0, r1), sum_r);
    sum_t = ffma2(valid, pk2(t0, t1), sum_t);
  }
  hr = cvt_h2<false>(vexp2n(r0), vexp2n(r1));
  ht = cvt_h2<false>(vexp2n(t0), vexp2n(t1));
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// One warp's share of a k-block of a hypothesis tile: 16 rows (row block rb) x 32 columns (column half chalf) = 4 groups
// of 8 columns; per group 4 HMMAs (u_x, u_y, u_z, t.u) and two row x column-pair tails per thread.
// Software-pipelined by hand: the HMMAs + FMA-pipe half of group g+1 are issued BEFORE the MUFU half of group g (both are
// volatile asm, so ptxas keeps that order), i.e. the 16 MUFU instructions of a group (128 cycles of the XU pipe) run under
// the next group's shared-memory loads, HMMA latency and packed-FMA chain instead of after them (r1j SASS: every group was
// one serial LDS -> HMMA -> FMA -> MUFU -> STS chain, 52 % of the MUFU floor).
// `sts0` = shared-window address of this thread's word in the rot A tile for group 0, row g: the 128-byte swizzle is
// ((chalf*4 + grp) ^ (row & 7)) << 4 and row & 7 == g for both rows, so group grp is sts0 ^ (grp << 4), row g+8 is +1024,
// the trans tile +BLK_BYTES.
template <bool SUMS>
__device__ __forceinline__ void residual_kblock_mma(const uint4 (&afrag)[4], const uint8_t* __restrict__ cjk, uint32_t sts0,
                                                    int chalf, int lane, int col0, int m, u64 (&sum_r)[2], u64 (&sum_t)[2]) {
  const int q = lane & 3;
  const uint2* bfr = reinterpret_cast<const uint2*>(cjk + 2048) + (chalf * 4) * 32 + lane;
  const ulonglong2* c8 = reinterpret_cast<const ulonglong2*>(cjk) + (chalf * 4) * (CJ8_FIELDS / 2 * 4) + q;
  u64 dr2[2][2], dt2[2][2];          // [pipeline slot][row half]
  auto front = [&](int grp, u64 (&r2)[2], u64 (&t2)[2]) {
    const uint2 b = bfr[grp * 32];
    float dx[4], dy[4], dz[4], dt[4];
#if NSAC_SCORE_ABLATE == 2
    for (int e = 0; e < 4; ++e) {
      dx[e] = __uint_as_float(afrag[0].x ^ b.x) * 1e-30f; dy[e] = __uint_as_float(afrag[1].y ^ b.y) * 1e-30f;
      dz[e] = __uint_as_float(afrag[2].z ^ b.x) * 1e-30f; dt[e] = __uint_as_float(afrag[3].w ^ b.y) * 1e-30f;
    }
#else
    hmma16816(dx, afrag[0], b.x, b.y);
    hmma16816(dy, afrag[1], b.x, b.y);
    hmma16816(dz, afrag[2], b.x, b.y);
    hmma16816(dt, afrag[3], b.x, b.y);
#endif
    u64 c[CJ8_FIELDS + 3];
#pragma unroll
    for (int i = 0; i < CJ8_FIELDS / 2; ++i) {
      const ulonglong2 v = c8[(grp * (CJ8_FIELDS / 2) + i) * 4];
      c[2 * i] = v.x; c[2 * i + 1] = v.y;
    }
    c[6] = fmul2(c[3], c[0]); c[7] = fmul2(c[3], c[1]); c[8] = fmul2(c[3], c[2]);          // -k pi1 = |pi1| * (-k n1)
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2)
      residual_dist2(pk2(dx[2 * h2], dx[2 * h2 + 1]), pk2(dy[2 * h2], dy[2 * h2 + 1]), pk2(dz[2 * h2], dz[2 * h2 + 1]),
                     pk2(dt[2 * h2], dt[2 * h2 + 1]), c, r2[h2], t2[h2]);
  };
  auto back = [&](int grp, const u64 (&r2)[2], const u64 (&t2)[2]) {
    u64 valid = 0ull;
    if (SUMS) {
      const int j = col0 + chalf * 32 + grp * 8 + 2 * q;
      valid = pk2(j < m ? 1.f : 0.f, j + 1 < m ? 1.f : 0.f);
    }
    const uint32_t a = sts0 ^ (uint32_t)(grp << 4);
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      uint32_t hr, ht;
      residual_finish<SUMS>(r2[h2], t2[h2], valid, hr, ht, sum_r[h2], sum_t[h2]);
      sts32(a + h2 * 1024, hr);
      sts32(a + h2 * 1024 + BLK_BYTES, ht);
    }
  };
#if NSAC_SCORE_SWPIPE
  front(0, dr2[0], dt2[0]);
#pragma unroll
  for (int grp = 0; grp < 4; ++grp) {
    if (grp + 1 < 4) front(grp + 1, dr2[(grp + 1) & 1], dt2[(grp + 1) & 1]);
    back(grp, dr2[grp & 1], dt2[grp & 1]);
  }
#else
#pragma unroll
  for (int grp = 0; grp < 4; ++grp) {
    front(grp, dr2[0], dt2[0]);
    back(grp, dr2[0], dt2[0]);
  }
#endif
}

// Optional in-kernel timeline (debug / profiling aid): one CTA records %globaltimer at role hand-offs into a host-provided
// buffer (nsac_debug_score_trace).  [role][event] = ns; roles: 0 residual, 1 mma, 2 epilogue, 3 gather; rows 4 / 5 = start /
// end time of every CTA.
constexpr int TRACE_EVENTS = 256;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define NSAC_TRACE(role, on)                                                                  \
  do {                                                                                        \
    if (p.trace && (int)blockIdx.x == p.trace_cta && (on) && trace_n < TRACE_EVENTS)                         \
      p.trace[(role) * TRACE_EVENTS + trace_n++] = gtime();                                   \
  } while (0)

struct TcParams {
  unsigned long long* trace;   // nullptr unless tracing
  int trace_cta;
  const float* geo_local;   // [B,NQ,6]
  const float* q_h;         // [B,NQ,4]
  const float* t_h;         // [B,NQ,3]
  const float* q0;          // [B,4]   hypothesis 0 = the initial pose
  const float* t0;          // [B,3]
  const float* feat_rot;    // [B,NQ,256]
  const float* feat_tran;
  const int32_t* matched_num;
  const float* vecs;        // packed: b1[2][128], b2[2][128], w34[2][128]
  const uint8_t* cjk;       // [B][NQp/64][4096 B] per k-block: column constants + HMMA B fragments of the hypothesis tiles
  const uint4* afr;         // [B][tiles*64 row pairs (g, g+8)][hi | lo][4 comps] uint4 = the HMMA A operand {a0,a1,a2,a3} as is
  int B, NQ, NQp, tiles_per_pair, need_sums;
  int num_items, row0_tiles, row0_at;     // item list = B*tiles_per_pair hypothesis tiles + row0_tiles inserted at index row0_at
  float* logits;            // [2][B][NQ+1]
  float* sums;              // [2][B][NQ+1]
  float* partials;          // [B][tiles][2][PART_STRIDE]
  // CVEC kernels: b1[2][128], b2[2][128], w34[2][128] as KERNEL PARAMETERS (constant bank): the epilogue warps read every one of
  // them once per hypothesis row and tile - as shared-memory broadcast loads that was 3072 of the ~17 K shared-memory
  // wavefronts per tile the kernel is bound by (r2a ncu: l1tex__data_pipe_lsu_wavefronts_mem_shared 49 % + TMA + UMMA reads)
  float cvec[6 * HID];
};

// Work items.  A hypothesis tile = (pair b, hypotheses 1 + 128*tile ... ) scored against the pair's m matched columns.
// A row-0 tile = hypothesis 0 (the initial pose, camera_head.py:991, 1019) of 128 consecutive pairs b .. b+127: row r
// belongs to pair b + r.  Its residual rows exp(-d) are one row per pair, so score_prep_kernel computes them next to the
// column constants ([2][B][NQp] fp16) and the TMA producer drops them straight into the A ring: the tile only costs its
// MMAs and epilogue and shares the whole barrier protocol (the residual warps just pass their A-stage turns).
struct Item { int b, tile, m, nkb; bool row0; };
__device__ __forceinline__ bool decode_item(const TcParams& p, int item, Item& it) {
  int idx = item;
  it.row0 = false;
  if (item >= p.row0_at) {
    if (item < p.row0_at + p.row0_tiles) {
      const int t = item - p.row0_at;
      it.row0 = true; it.b = t * TILE_H; it.tile = 0; it.m = 0; it.nkb = p.NQp / KB;
      return true;
    }
    idx = item - p.row0_tiles;
  }
  it.b = idx / p.tiles_per_pair; it.tile = idx - it.b * p.tiles_per_pair;
  it.m = p.matched_num[it.b];
  it.nkb = (it.m + KB - 1) / KB;
  return it.tile * TILE_H < it.m;
}

template <bool SUMS, bool CVEC>
__global__ void __launch_bounds__(NUM_THREADS, 1)
score_tc_kernel(const __grid_constant__ CUtensorMap map_w1r, const __grid_constant__ CUtensorMap map_w1t,
                const __grid_constant__ CUtensorMap map_w2r, const __grid_constant__ CUtensorMap map_w2t,
                const __grid_constant__ CUtensorMap map_x0r, const __grid_constant__ CUtensorMap map_x0t, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // align inside the shared window with offset arithmetic (a generic-pointer round trip would turn every
  // shared-memory access below into a generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  float* vec = reinterpret_cast<float*>(smem + OFF_VEC);
  float* s_logit = reinterpret_cast<float*>(smem + OFF_LOGIT);
  float* s_rowsum = reinterpret_cast<float*>(smem + OFF_ROWSUM);
  float* s_exp = reinterpret_cast<float*>(smem + OFF_EXP);
  float* s_gpart = reinterpret_cast<float*>(smem + OFF_GPART);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H1n = p.NQ + 1;
  const int num_items = p.num_items;
  int trace_n = 0;
  if (p.trace && threadIdx.x == 0 && blockIdx.x < TRACE_EVENTS) p.trace[4 * TRACE_EVENTS + blockIdx.x] = gtime();     // CTA start

  if (threadIdx.x == 0) {
    for (int s = 0; s < W_STAGES; ++s) { mbar_init(&bars[BAR_W_FULL + s], 1); mbar_init(&bars[BAR_W_EMPTY + s], 1); }
    for (int s = 0; s < A_STAGES; ++s) { mbar_init(&bars[BAR_CJ_FULL + s], 1); mbar_init(&bars[BAR_CJ_EMPTY + s], R_WARPS); }
    for (int s = 0; s < F_STAGES; ++s) { mbar_init(&bars[BAR_F_FULL + s], 1); mbar_init(&bars[BAR_F_EMPTY + s], G_THREADS / 64); }
    for (int s = 0; s < A_STAGES; ++s) { mbar_init(&bars[BAR_A_FULL + s], R_WARPS + 1); mbar_init(&bars[BAR_A_EMPTY + s], 1); }
    mbar_init(&bars[BAR_ACC_FULL], 1);
    mbar_init(&bars[BAR_H1_READY], 4);
    mbar_init(&bars[BAR_ACC2_FULL], 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&bars[BAR_LOGIT_READY + s], 4); mbar_init(&bars[BAR_LOGIT_FREE + s], G_THREADS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (!CVEC)
    for (int i = threadIdx.x; i < 6 * HID; i += NUM_THREADS) vec[i] = p.vecs[i];
  if (warp == M_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above (barriers, tensor memory, the weight vectors) is independent of score_prep_kernel: with programmatic
  // dependent launch it overlaps the prep kernel's tail.  From here on its outputs (cjk, afr, x0) are read.
  pdl_wait();
  pdl_launch_dependents();          // the selection kernel may be scheduled as soon as SMs free up (it waits for this grid)

  // setmaxnreg: ONE instruction per warpgroup (all 4 warps must execute the same one), inside the warpgroup's own
  // branch so that the register limit is unambiguous on every control path
  if (warp >= T_WARP) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == T_WARP) {
    // ================================================================================= weight producer (W ring)
    // per tile: nkb stages of W1 k-blocks, then two stages carrying the tile's W2 k-blocks (layer 2)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        Item it;
        if (!decode_item(p, item, it)) continue;
        for (int kb = 0; kb < it.nkb + 2; ++kb) {
          mbar_wait(&bars[BAR_W_EMPTY + stage], phase ^ 1);
          mbar_expect_tx(&bars[BAR_W_FULL + stage], 2 * BLK_BYTES);
          uint8_t* dst = smem + OFF_W1 + stage * 2 * BLK_BYTES;
          if (kb < it.nkb) {
            tma_load_2d(dst, &map_w1r, &bars[BAR_W_FULL + stage], kb * KB, 0);
            tma_load_2d(dst + BLK_BYTES, &map_w1t, &bars[BAR_W_FULL + stage], kb * KB, 0);
          } else {
            tma_load_2d(dst, &map_w2r, &bars[BAR_W_FULL + stage], (kb - it.nkb) * KB, 0);
            tma_load_2d(dst + BLK_BYTES, &map_w2t, &bars[BAR_W_FULL + stage], (kb - it.nkb) * KB, 0);
          }
          if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == C_WARP) {
    // ================================================================================= column-block / A-ring producer
    // hypothesis tiles: the k-block's column block (2 KB of constants + 2 KB of HMMA B fragments) into the CJ ring, and the
    // 17th arrival on the A stage; row-0 tiles: that arrival carries the two [128 pairs x 64 columns] fp16 residual
    // blocks TMA writes into the A stage (rows >= B: zero fill)
    if (lane == 0) {
      int as = 0; uint32_t aph = 0; int cs = 0; uint32_t cph = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        Item it;
        if (!decode_item(p, item, it)) continue;
        for (int kb = 0; kb < it.nkb; ++kb) {
          if (!it.row0) {
            mbar_wait(&bars[BAR_CJ_EMPTY + cs], cph ^ 1);
            mbar_expect_tx(&bars[BAR_CJ_FULL + cs], CJK_BYTES);
            bulk_load_1d(smem + OFF_CJ + cs * CJK_BYTES, p.cjk + ((size_t)it.b * (p.NQp / KB) + kb) * CJK_BYTES, CJK_BYTES,
                         &bars[BAR_CJ_FULL + cs]);
            if (++cs == A_STAGES) { cs = 0; cph ^= 1; }
          }
          mbar_wait(&bars[BAR_A_EMPTY + as], aph ^ 1);
          if (it.row0) {
            uint8_t* a = smem + OFF_A + as * 2 * BLK_BYTES;
            mbar_expect_tx(&bars[BAR_A_FULL + as], 2 * BLK_BYTES);
            tma_load_2d(a, &map_x0r, &bars[BAR_A_FULL + as], kb * KB, it.b);
            tma_load_2d(a + BLK_BYTES, &map_x0t, &bars[BAR_A_FULL + as], kb * KB, it.b);
          } else {
            mbar_arrive(&bars[BAR_A_FULL + as]);
          }
          if (++as == A_STAGES) { as = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == F_WARP) {
    // ================================================================================= feature producer (F ring)
    // The only HBM stream of the kernel.  256 gather threads with register-staged LDG.128 top out at ~3.6 TB/s no matter
    // how many loads are in flight; cp.async.bulk into a 8 x 8 KB ring reaches 6.2 TB/s with 64 KB in flight (scripts/ubench/stream.cu).  The
    // producer runs ahead of the consumers across tile boundaries: chunk = F_ROWS hypothesis rows of one branch, order
    // (chunk 0, rot), (chunk 0, tran), (chunk 1, rot), ...
    if (lane == 0) {
      int fs = 0; uint32_t fph = 0;
#if NSAC_SCORE_L2PREFETCH
      // Experiment (r2b, rejected): pulling the NEXT tile's rows into L2 one item ahead (cp.async.bulk.prefetch.L2) so that HBM
      // streams during the first tile's scoring.  Measured: DRAM reads 286 -> 393 MB per launch (prefetched lines are evicted
      // or fetched twice) and the kernel got 10 us slower; the ring fills were no faster from L2.  Kept behind the macro.
      auto prefetch_item = [&](int item) {
        Item nx;
        if (item >= num_items || !decode_item(p, item, nx) || nx.row0) return;
        const uint32_t bytes = (uint32_t)min(TILE_H, nx.m - nx.tile * TILE_H) * (C_FEAT * 4);
        const size_t off = ((size_t)nx.b * p.NQ + (size_t)nx.tile * TILE_H) * C_FEAT;
        for (uint32_t o = 0; o < bytes; o += 32768u) {
          const uint32_t n = min(32768u, bytes - o);
          l2_prefetch(reinterpret_cast<const uint8_t*>(p.feat_rot + off) + o, n);
          l2_prefetch(reinterpret_cast<const uint8_t*>(p.feat_tran + off) + o, n);
        }
      };
      prefetch_item(blockIdx.x);
#else
      auto prefetch_item = [&](int) {};
#endif
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        Item it;
        prefetch_item(item + gridDim.x);
        if (!decode_item(p, item, it) || it.row0) continue;
        const int rows = min(TILE_H, it.m - it.tile * TILE_H);
        const size_t row0 = (size_t)it.b * p.NQ + (size_t)it.tile * TILE_H;
        for (int r = 0; r < rows; r += F_ROWS) {
          const uint32_t bytes = (uint32_t)min(F_ROWS, rows - r) * (C_FEAT * 4);
#pragma unroll 1
          for (int br = 0; br < 2; ++br) {
            mbar_wait(&bars[BAR_F_EMPTY + fs], fph ^ 1);
            NSAC_TRACE(6, true);
            mbar_expect_tx(&bars[BAR_F_FULL + fs], bytes);
            bulk_load_1d(smem + OFF_F + fs * F_BYTES, (br == 0 ? p.feat_rot : p.feat_tran) + (row0 + r) * C_FEAT, bytes, &bars[BAR_F_FULL + fs]);
            if (++fs == F_STAGES) { fs = 0; fph ^= 1; }
          }
        }
      }
    }
  } else if (warp == M_WARP) {
    // ================================================================================= MMA issuer
    if (lane == 0) {
      int ws = 0; uint32_t wph = 0; int as = 0; uint32_t aph = 0; uint32_t tph = 0;   // tph: per-tile barrier parity
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        Item it;
        if (!decode_item(p, item, it)) continue;
        // layer 1 may overwrite D / H1 of the previous tile: its layer-2 MMAs (the last readers) were issued earlier
        // and tcgen05.mma executes in issue order; the epilogue warps finished reading D before they signalled H1_READY
        for (int kb = 0; kb < it.nkb; ++kb) {
          mbar_wait(&bars[BAR_W_FULL + ws], wph);
          NSAC_TRACE(1, true);
          mbar_wait(&bars[BAR_A_FULL + as], aph);
          NSAC_TRACE(1, true);
          fence_after();
          const uint32_t a0 = smem_u32(smem + OFF_A + as * 2 * BLK_BYTES), w0 = smem_u32(smem + OFF_W1 + ws * 2 * BLK_BYTES);
#pragma unroll 1
          for (int br = 0; br < 2; ++br) {
            const uint64_t da = sw128_desc(a0 + br * BLK_BYTES), dw = sw128_desc(w0 + br * BLK_BYTES);
#pragma unroll 1
            for (int k = 0; k < KB / 16; ++k)
              if (NSAC_SCORE_ABLATE != 5) umma_ss(tmem_base + TM_D + br * HID, da + 2 * k, dw + 2 * k, IDESC, (kb | k) != 0);
          }
          umma_commit(&bars[BAR_W_EMPTY + ws]);
          umma_commit(&bars[BAR_A_EMPTY + as]);
          if (++ws == W_STAGES) { ws = 0; wph ^= 1; }
          if (++as == A_STAGES) { as = 0; aph ^= 1; }
        }
        umma_commit(&bars[BAR_ACC_FULL]);
        // layer 2: A = H1 (fp16 pairs in tensor memory), B = W2 (two k-blocks through the W ring), D2 in its own columns.
        // H1_READY of this tile also implies that the epilogue warps have finished reading D2 of the previous tile.
        NSAC_TRACE(1, true);
        mbar_wait(&bars[BAR_H1_READY], tph);
        NSAC_TRACE(1, true);
        fence_after();
#pragma unroll 1
        for (int kb2 = 0; kb2 < 2; ++kb2) {
          mbar_wait(&bars[BAR_W_FULL + ws], wph);
          fence_after();
          const uint32_t w0 = smem_u32(smem + OFF_W1 + ws * 2 * BLK_BYTES);
#pragma unroll 1
          for (int br = 0; br < 2; ++br) {
            const uint64_t dw = sw128_desc(w0 + br * BLK_BYTES);
#pragma unroll 1
            for (int k4 = 0; k4 < 4; ++k4)
              if (NSAC_SCORE_ABLATE != 5) umma_ts(tmem_base + TM_D2 + br * HID, tmem_base + TM_D + br * HID + (kb2 * 4 + k4) * 8, dw + 2 * k4, IDESC, (kb2 | k4) != 0);
          }
          umma_commit(&bars[BAR_W_EMPTY + ws]);
          if (++ws == W_STAGES) { ws = 0; wph ^= 1; }
        }
        umma_commit(&bars[BAR_ACC2_FULL]);
        tph ^= 1;
      }
    }
  }
  } else if (warp >= E_WARP0) {
    // ================================================================================= epilogue warps 24..27
    asm volatile("setmaxnreg.inc.sync.aligned.u32 72;");
    const int quad = warp & 3, row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    uint32_t tph = 0; int lbuf = 0; uint32_t lph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      Item it;
      if (!decode_item(p, item, it)) continue;
      // ---- layer-1 epilogue: H1 = fp16(relu(D + b1)) -> tensor memory, in place over the columns already read
      NSAC_TRACE(2, threadIdx.x == E_WARP0 * 32);
      mbar_wait(&bars[BAR_ACC_FULL], tph);
      NSAC_TRACE(2, threadIdx.x == E_WARP0 * 32);
      fence_after();
#pragma unroll
      for (int br = 0; br < 2; ++br) {
        const float* b1 = (CVEC ? p.cvec : vec) + br * HID;
#pragma unroll
        for (int c = 0; c < HID; c += 32) {
          uint32_t v[32], h[16];
          tmem_ld32(tmem_base + lane_base + TM_D + br * HID + c, v);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {       // packed bias add, ReLU inside the fp16x2 conversion
            float lo, hi;
            upk2(fadd2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), *reinterpret_cast<const u64*>(b1 + c + i)), lo, hi);
            h[i >> 1] = cvt_h2<true>(lo, hi);
          }
          tmem_st16(tmem_base + lane_base + TM_D + br * HID + (c >> 1), h);
        }
      }
      tmem_st_wait();
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BAR_H1_READY]);
      NSAC_TRACE(2, threadIdx.x == E_WARP0 * 32);
      // ---- layer-2 epilogue: logit = w34 . relu(D2 + b2) + c34   (this thread owns hypothesis row `row`)
      mbar_wait(&bars[BAR_ACC2_FULL], tph);
      NSAC_TRACE(2, threadIdx.x == E_WARP0 * 32);
      fence_after();
      float lg[2];
#pragma unroll
      for (int br = 0; br < 2; ++br) {
        const float* b2 = (CVEC ? p.cvec : vec) + 2 * HID + br * HID;
        const float* w34 = (CVEC ? p.cvec : vec) + 4 * HID + br * HID;
        u64 acc2[2] = {0ull, 0ull};
#pragma unroll
        for (int c = 0; c < HID; c += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + lane_base + TM_D2 + br * HID + c, v);
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float lo, hi;
            upk2(fadd2(pk2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), *reinterpret_cast<const u64*>(b2 + c + i)), lo, hi);
            acc2[(i >> 1) & 1] = ffma2(pk2(fmaxf(lo, 0.f), fmaxf(hi, 0.f)), *reinterpret_cast<const u64*>(w34 + c + i), acc2[(i >> 1) & 1]);
          }
        }
        float s0, s1, s2, s3;
        upk2(acc2[0], s0, s1);
        upk2(acc2[1], s2, s3);
        lg[br] = (s0 + s1) + (s2 + s3);      // the constant c34 cancels in the softmax; it is added by the selection kernel
      }
      fence_before();
      NSAC_TRACE(2, threadIdx.x == E_WARP0 * 32);
      // ---- hand the logits to the gather warps (and to global memory for scores / argmax)
      mbar_wait(&bars[BAR_LOGIT_FREE + lbuf], lph ^ 1);
      NSAC_TRACE(2, threadIdx.x == E_WARP0 * 32);
      s_logit[(lbuf * 2 + 0) * TILE_H + row] = lg[0];
      s_logit[(lbuf * 2 + 1) * TILE_H + row] = lg[1];
      if (it.row0) {                          // row r of a row-0 tile = hypothesis 0 of pair b + r
        if (it.b + row < p.B) {
          p.logits[(size_t)(it.b + row) * H1n] = lg[0];
          p.logits[(size_t)p.B * H1n + (size_t)(it.b + row) * H1n] = lg[1];
        }
      } else {
        const int h = 1 + it.tile * TILE_H + row;
        if (h <= it.m) {
          p.logits[(size_t)it.b * H1n + h] = lg[0];
          p.logits[(size_t)p.B * H1n + (size_t)it.b * H1n + h] = lg[1];
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BAR_LOGIT_READY + lbuf]);
      if (++lbuf == 2) { lbuf = 0; lph ^= 1; }
      tph ^= 1;
    }
  } else if (warp >= R_WARP0) {
    // ================================================================================= residual warps 8..23
    // Hypothesis tiles: warp rw = row block rw % 8 (16 rows) x column half rw / 8 (32 columns) per k-block; u and t.u come
    // from 4 HMMAs per 8 columns (residual_kblock_mma), the rest (13 packed FMA-pipe instructions + 4 MUFU.SQRT + 4
    // MUFU.EX2 + 2 cvt per row x column pair) runs on the CUDA cores.  Row-0 tiles: the A stages are filled by TMA, these
    // warps only pass their turns.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 80;");
    const int rw = warp - R_WARP0;                   // 0..15
    const int rb = rw & 7, chalf = rw >> 3, g = lane >> 2, q = lane & 3;
    // this thread's word of row rb*16 + g, group 0 of its column half, in A stage 0 (128-byte swizzle; see residual_kblock_mma)
    const uint32_t sts_base = smem_u32(smem + OFF_A) + (uint32_t)(rb * 16 + g) * 128u + (uint32_t)(((chalf * 4) ^ g) << 4) + (uint32_t)(q << 2);
    int as = 0; uint32_t aph = 0; int cs = 0; uint32_t cph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      Item it;
      if (!decode_item(p, item, it)) continue;
      if (it.row0) {
        for (int kb = 0; kb < it.nkb; ++kb) {
          mbar_wait(&bars[BAR_A_EMPTY + as], aph ^ 1);
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[BAR_A_FULL + as]);
          if (++as == A_STAGES) { as = 0; aph ^= 1; }
        }
        continue;
      }
      // ---- hypothesis tile: A fragments of this thread's two rows (g, g + 8 of row block rb), once per tile.  They were
      // computed by score_prep_kernel for every hypothesis (identity pose beyond m): two 16-byte loads per row, no math
      // and no dependence on matched_num, instead of two dependent global round trips + ~250 instructions per tile.
      uint4 afrag[4];       // one HMMA A operand (a0..a3) per component: u_x, u_y, u_z, t.u
      {
        const uint4* ar = p.afr + (((size_t)it.b * p.tiles_per_pair + it.tile) * (TILE_H / 2) + rb * 8 + g) * 8 + (q == 1 ? 4 : 0);
#pragma unroll
        for (int c = 0; c < 4; ++c) afrag[c] = q == 3 ? make_uint4(0u, 0u, 0u, 0u) : __ldg(ar + c);
      }
      u64 sum_r[2] = {0ull, 0ull}, sum_t[2] = {0ull, 0ull};
      // decode the next item now and prefetch its fragment rows after the first k-block
      Item nx;
      nx.row0 = true;
      if (item + (int)gridDim.x < num_items && !decode_item(p, item + gridDim.x, nx)) nx.row0 = true;
      for (int kb = 0; kb < it.nkb; ++kb) {
        if (kb == 1 && !nx.row0 && q == 0) {        // next tile's fragments of this row pair (128 B) into L2
          const uint4* ar = p.afr + (((size_t)nx.b * p.tiles_per_pair + nx.tile) * (TILE_H / 2) + rb * 8 + g) * 8;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(ar));
        }
        NSAC_TRACE(0, threadIdx.x == R_WARP0 * 32);
        mbar_wait(&bars[BAR_A_EMPTY + as], aph ^ 1);
        NSAC_TRACE(0, threadIdx.x == R_WARP0 * 32);
        mbar_wait(&bars[BAR_CJ_FULL + cs], cph);          // column block of this k-block has landed (TMA)
        NSAC_TRACE(0, threadIdx.x == R_WARP0 * 32);
#if NSAC_SCORE_ABLATE != 3
        residual_kblock_mma<SUMS>(afrag, smem + OFF_CJ + cs * CJK_BYTES, sts_base + (uint32_t)(as * 2 * BLK_BYTES), chalf, lane, kb * KB,
                                  it.m, sum_r, sum_t);
#endif
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&bars[BAR_A_FULL + as]); mbar_arrive(&bars[BAR_CJ_EMPTY + cs]); }
        if (++as == A_STAGES) { as = 0; aph ^= 1; }
        if (++cs == A_STAGES) { cs = 0; cph ^= 1; }
      }
      if (SUMS) {   // sum_j of the masked distances (argmin in 'min-cost', :1090-1093): 4 q-lanes x 2 column halves per row
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          float a0, a1, c0, c1;
          upk2(sum_r[h2], a0, a1);
          upk2(sum_t[h2], c0, c1);
          float a = a0 + a1, c = c0 + c1;
          a += __shfl_xor_sync(NSAC_FULL_MASK, a, 1); a += __shfl_xor_sync(NSAC_FULL_MASK, a, 2);
          c += __shfl_xor_sync(NSAC_FULL_MASK, c, 1); c += __shfl_xor_sync(NSAC_FULL_MASK, c, 2);
          if (q == 0) {
            s_rowsum[(chalf * 2 + 0) * TILE_H + rb * 16 + g + 8 * h2] = a * (1.f / LOG2E);
            s_rowsum[(chalf * 2 + 1) * TILE_H + rb * 16 + g + 8 * h2] = c * (1.f / LOG2E);
          }
        }
        named_bar_sync(1, R_THREADS);
        const int rt = threadIdx.x - R_WARP0 * 32;
        if (rt < TILE_H && it.tile * TILE_H + rt < it.m) {
          const int h = it.tile * TILE_H + rt + 1;
          p.sums[(size_t)it.b * H1n + h] = s_rowsum[0 * TILE_H + rt] + s_rowsum[2 * TILE_H + rt];
          p.sums[(size_t)p.B * H1n + (size_t)it.b * H1n + h] = s_rowsum[1 * TILE_H + rt] + s_rowsum[3 * TILE_H + rt];
        }
        named_bar_sync(1, R_THREADS);
      }
    }
  } else {
    // ================================================================================= gather warps 0..7
    // consumers of the feature ring (F_WARP): flash-style softmax partials of the tile (local max / sum of exp, the
    // exp-weighted and the plain sum of its [rows,256] one-plane features)
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    const int gt = threadIdx.x - G_WARP0 * 32;       // 0..255
    const int br = gt >> 7, rs = (gt >> 6) & 1, t64 = gt & 63, c4 = t64 * 4;
    int lbuf = 0; uint32_t lph = 0;
    int fs = br; uint32_t fph = 0;                   // this branch's next stage of the feature ring (stages alternate rot / tran)
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      Item it;
      if (!decode_item(p, item, it)) continue;
      NSAC_TRACE(3, threadIdx.x == G_WARP0 * 32);
      mbar_wait(&bars[BAR_LOGIT_READY + lbuf], lph);
      NSAC_TRACE(3, threadIdx.x == G_WARP0 * 32);
      if (it.row0) {                          // hypothesis 0 joins in the selection kernel: nothing to gather
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BAR_LOGIT_FREE + lbuf]);
        if (++lbuf == 2) { lbuf = 0; lph ^= 1; }
        continue;
      }
      const int b = it.b, tile = it.tile;
      const int rows = min(TILE_H, it.m - tile * TILE_H);
      const float* lg = s_logit + (lbuf * 2 + br) * TILE_H;
      // tile maximum: own row -> warp shuffle -> 4 per-warp maxima through shared memory (slot 9 of s_gpart rows 0..3)
      float* ex = s_exp + br * TILE_H;
      float mx;
      {
        const int r = rs * 64 + t64;
        const float v = r < rows ? lg[r] : -INFINITY;
        const float wm = warp_max(v);
        float* wmax = s_gpart + (br * 64) * 12 + 9;
        if (lane == 0) wmax[((gt >> 5) & 3) * 12] = wm;
        named_bar_sync(4 + br, 128);
        mx = fmaxf(fmaxf(wmax[0], wmax[12]), fmaxf(wmax[24], wmax[36]));
        ex[r] = r < rows ? __expf(v - mx) : 0.f;
      }
      named_bar_sync(4 + br, 128);
      // packed accumulators: (x,y) and (z,w) of the exp-weighted sum and of the plain sum
      u64 wxy = 0ull, wzw = 0ull, fxy = 0ull, fzw = 0ull;
      float se = 0.f;
      if (rs == 0 && t64 < 32) {             // sum of the tile's softmax numerators, once (one warp per branch)
        const float4 e4 = *reinterpret_cast<const float4*>(ex + t64 * 4);
        se = warp_sum((e4.x + e4.y) + (e4.z + e4.w));
      }
      // feature ring: this thread = 4 channels x rows rs*8 .. rs*8+7 of every 16-row chunk of its branch
      constexpr int HR = F_ROWS / 2;
      for (int r0 = 0; r0 < rows; r0 += F_ROWS) {
        mbar_wait(&bars[BAR_F_FULL + fs], fph);
        NSAC_TRACE(7, threadIdx.x == G_WARP0 * 32);
        const ulonglong2* fr = reinterpret_cast<const ulonglong2*>(smem + OFF_F + fs * F_BYTES) + (rs * HR) * (C_FEAT / 4) + t64;
        float e[HR];
#pragma unroll
        for (int i = 0; i < HR; i += 4) {
          const float4 e4 = *reinterpret_cast<const float4*>(ex + r0 + rs * HR + i);      // 0 beyond `rows`
          e[i] = e4.x; e[i + 1] = e4.y; e[i + 2] = e4.z; e[i + 3] = e4.w;
        }
#pragma unroll
        for (int i = 0; i < HR; ++i) {
          if (NSAC_SCORE_ABLATE != 4 && r0 + rs * HR + i < rows) {            // (a partial last chunk leaves stale rows in the stage)
            const ulonglong2 v = fr[i * (C_FEAT / 4)];
            const u64 e2 = bc2(e[i]);
            wxy = ffma2(e2, v.x, wxy); wzw = ffma2(e2, v.y, wzw);
            fxy = fadd2(fxy, v.x); fzw = fadd2(fzw, v.y);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[BAR_F_EMPTY + fs]);
        fs += 2;
        if (fs >= F_STAGES) { fs -= F_STAGES; fph ^= 1; }
      }
      float4 ws, fs4;
      upk2(wxy, ws.x, ws.y); upk2(wzw, ws.z, ws.w);
      upk2(fxy, fs4.x, fs4.y); upk2(fzw, fs4.z, fs4.w);
      float* gp = s_gpart + (br * 64 + t64) * 12;
      if (rs == 1) {
        *reinterpret_cast<float4*>(gp) = ws;
        *reinterpret_cast<float4*>(gp + 4) = fs4;
      }
      named_bar_sync(4 + br, 128);
      if (rs == 0) {
        const float4 w1 = *reinterpret_cast<const float4*>(gp), f1 = *reinterpret_cast<const float4*>(gp + 4);
        ws.x += w1.x; ws.y += w1.y; ws.z += w1.z; ws.w += w1.w;
        fs4.x += f1.x; fs4.y += f1.y; fs4.z += f1.z; fs4.w += f1.w;
        float* part = p.partials + (((size_t)b * p.tiles_per_pair + tile) * 2 + br) * PART_STRIDE;
        if (t64 == 0) { part[0] = mx; part[1] = se; }
        *reinterpret_cast<float4*>(part + PART_HDR + c4) = ws;
        *reinterpret_cast<float4*>(part + PART_HDR + C_FEAT + c4) = fs4;
      }
      named_bar_sync(4 + br, 128);      // s_exp / s_gpart free for the next tile
      NSAC_TRACE(3, threadIdx.x == G_WARP0 * 32);
      if (lane == 0) mbar_arrive(&bars[BAR_LOGIT_FREE + lbuf]);
      if (++lbuf == 2) { lbuf = 0; lph ^= 1; }
    }
  }
  fence_before();
  __syncthreads();
  if (p.trace && threadIdx.x == 0 && blockIdx.x < TRACE_EVENTS) p.trace[5 * TRACE_EVENTS + blockIdx.x] = gtime();     // CTA end
  if (warp == M_WARP) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TM_COLS));
  }
}

// Two blocks per pair: the A fragments of its hypothesis rows / the column blocks of its hypothesis tiles and hypothesis 0
// (the initial pose).
//   cjk [B][NQp/64][4096 B]    per k-block: column constants -k n1, -k pi1, A, Bc (CJ8 layout) + HMMA B fragments
//                              [8 groups][32 lanes][2 words] of n^ (fp16 hi / hi / lo k-slots, see hmma16816)
//   x0  [2][B][NQp] fp16       exp(-d) of hypothesis 0 against every column (0 beyond m): the A operand of the row-0 tiles
//   sums[.][b][0]              masked distance sums of hypothesis 0 ('min-cost'), fixed summation order
constexpr int PREP_THREADS = 256;
__global__ void __launch_bounds__(PREP_THREADS)
score_prep_kernel(const float* __restrict__ geo_local, const float* __restrict__ q_h, const float* __restrict__ t_h,
                  const float* __restrict__ q0, const float* __restrict__ t0, const int32_t* __restrict__ matched_num, int B, int NQ,
                  int NQp, int rows_per_pair, uint8_t* __restrict__ cjk, __half* __restrict__ x0, uint4* __restrict__ afr,
                  float* __restrict__ sums) {
  __shared__ float red[2][PREP_THREADS / 32];
  pdl_launch_dependents();          // the tile kernel's prologue (barriers, tensor memory) may overlap this kernel
  // blocks [0, B): the rows of pair b (A fragments); blocks [B, 2B): its columns + hypothesis 0 - two independent load
  // chains, so they run as separate blocks instead of back to back
  const bool do_rows = (int)blockIdx.x < B;
  const int b = do_rows ? blockIdx.x : blockIdx.x - B, tid = threadIdx.x, m = matched_num[b];
  // A fragments of every hypothesis row h = 1 + i (rows beyond m: identity pose; their D rows are never read).  One thread per
  // row PAIR (g, g + 8) of a 16-row block: the pair's fragments are stored as the four HMMA A operands {a0 = row g k-slots
  // 2q,2q+1 | a1 = row g+8 | a2 = row g k-slots 2q+8,2q+9 | a3 = row g+8}, hi words (lanes q = 0, 2) then lo words (q = 1), so the
  // residual warps load them straight into the operand registers.
  for (int pi = tid; do_rows && pi < rows_per_pair / 2; pi += PREP_THREADS) {
    uint32_t w01[2][2][4], w2[2][2][4];          // [row of the pair][hi / lo][component]
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int i = (pi >> 3) * 16 + (pi & 7) + 8 * rr;
      float qw = 1.f, qx = 0.f, qy = 0.f, qz = 0.f, tx = 0.f, ty = 0.f, tz = 0.f;
      if (i < m) {
        const float4 qq = *reinterpret_cast<const float4*>(q_h + ((size_t)b * NQ + i) * 4);
        const float* tp = t_h + ((size_t)b * NQ + i) * 3;
        qw = qq.x; qx = qq.y; qy = qq.z; qz = qq.w;
        tx = tp[0]; ty = tp[1]; tz = tp[2];
      }
      float Rh[9], th[3];
      row_constants(qw, qx, qy, qz, tx, ty, tz, Rh, th);
      // t.u = (t / k) . ((k R) n^) = s . n^ with s = (k R)^T (t / k)
      const float sv[3] = {fmaf(th[0], Rh[0], fmaf(th[1], Rh[3], th[2] * Rh[6])), fmaf(th[0], Rh[1], fmaf(th[1], Rh[4], th[2] * Rh[7])),
                           fmaf(th[0], Rh[2], fmaf(th[1], Rh[5], th[2] * Rh[8]))};
#pragma unroll
      for (int v = 0; v < 2; ++v) {        // v = 0: hi words (k-slot groups q = 0, 2), v = 1: lo words (q = 1)
        a_words(Rh[0], Rh[1], Rh[2], v, w01[rr][v][0], w2[rr][v][0]);
        a_words(Rh[3], Rh[4], Rh[5], v, w01[rr][v][1], w2[rr][v][1]);
        a_words(Rh[6], Rh[7], Rh[8], v, w01[rr][v][2], w2[rr][v][2]);
        a_words(sv[0], sv[1], sv[2], v, w01[rr][v][3], w2[rr][v][3]);
      }
    }
    uint4* o = afr + ((size_t)b * (rows_per_pair / 2) + pi) * 8;
#pragma unroll
    for (int v = 0; v < 2; ++v)
#pragma unroll
      for (int c = 0; c < 4; ++c) o[v * 4 + c] = make_uint4(w01[0][v][c], w01[1][v][c], w2[0][v][c], w2[1][v][c]);
  }
  if (do_rows) return;
  float R[9], ts[3];
  row_constants(q0[b * 4 + 0], q0[b * 4 + 1], q0[b * 4 + 2], q0[b * 4 + 3], t0[b * 3 + 0], t0[b * 3 + 1], t0[b * 3 + 2], R, ts);
  float sr = 0.f, st = 0.f;
  for (int j = tid; j < NQp; j += PREP_THREADS) {
    float f[CJ_FIELDS];
    const bool valid = j < m;
    column_fields(geo_local + ((size_t)b * NQ + (valid ? j : 0)) * 6, valid, f);
    uint8_t* blk = cjk + ((size_t)b * (NQp / KB) + j / KB) * CJK_BYTES;
    const int jj = j % KB;
    // pair jj/2 = group (jj/8) of 4 pairs, q = (jj/2) % 4; field k -> unit k/2, field parity k%2
    float* c8 = reinterpret_cast<float*>(blk) + (jj >> 3) * (CJ8_FIELDS / 2 * 16) + ((jj >> 1) & 3) * 4 + (jj & 1);
    const float f6[CJ8_FIELDS] = {f[3], f[4], f[5], f[12], f[9], f[10]};
#pragma unroll
    for (int k = 0; k < CJ8_FIELDS; ++k) c8[(k >> 1) * 16 + (k & 1) * 2] = f6[k];
    // B fragments: column jj -> group jj / 8, g = jj % 8; lanes 4g + q hold (k = 2q, 2q+1 | 2q+8, 2q+9)
    uint32_t hi01, hi2, lo01, lo2;
    a_words(f[0], f[1], f[2], 0, hi01, hi2);
    a_words(f[0], f[1], f[2], 1, lo01, lo2);
    uint2* bf = reinterpret_cast<uint2*>(blk + 2048) + (jj >> 3) * 32 + (jj & 7) * 4;
    bf[0] = make_uint2(hi01, hi2);
    bf[1] = make_uint2(hi01, hi2);
    bf[2] = make_uint2(lo01, lo2);
    bf[3] = make_uint2(0u, 0u);
    // hypothesis 0 against column j
    float xr, xt;
    residual_scalar<true>(R, ts, f, xr, xt, sr, st);
    x0[(size_t)b * NQp + j] = __float2half_rn(valid ? xr : 0.f);
    x0[(size_t)B * NQp + (size_t)b * NQp + j] = __float2half_rn(valid ? xt : 0.f);
  }
  sr = warp_sum(sr); st = warp_sum(st);
  if ((tid & 31) == 0) { red[0][tid >> 5] = sr; red[1][tid >> 5] = st; }
  __syncthreads();
  if (tid == 0) {
    float a = 0.f, c = 0.f;
    for (int w = 0; w < PREP_THREADS / 32; ++w) { a += red[0][w]; c += red[1][w]; }
    sums[(size_t)b * (NQ + 1)] = a * (1.f / LOG2E);
    sums[(size_t)B * (NQ + 1) + (size_t)b * (NQ + 1)] = c * (1.f / LOG2E);
  }
}

// ------------------------------------------------------------------------------------------------ weight packing
// pack layout: [w1h_rot fp16 128 x NQp][w1h_tran][w2h_rot fp16 128x128][w2h_tran]
//              [vecs fp32: b1[2][128] b2[2][128] w34[2][128]][c34[2] + pad to 8]
__host__ __device__ inline size_t pack_w1_bytes(int NQp) { return (size_t)HID * NQp * 2; }
__host__ __device__ inline size_t pack_off_vecs(int NQp) { return 2 * pack_w1_bytes(NQp) + 2 * (size_t)HID * HID * 2; }
__host__ __device__ inline size_t pack_total_bytes(int NQp) { return pack_off_vecs(NQp) + (6 * HID + 8) * sizeof(float); }

__global__ void score_pack_kernel(nsac_score_mlp r, nsac_score_mlp t, int NQ, int NQp, uint8_t* pack) {
  __half* w1[2] = {reinterpret_cast<__half*>(pack), reinterpret_cast<__half*>(pack + pack_w1_bytes(NQp))};
  __half* w2[2] = {reinterpret_cast<__half*>(pack + 2 * pack_w1_bytes(NQp)),
                   reinterpret_cast<__half*>(pack + 2 * pack_w1_bytes(NQp) + HID * HID * 2)};
  float* vecs = reinterpret_cast<float*>(pack + pack_off_vecs(NQp));
  float* c34 = vecs + 6 * HID;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int br = 0; br < 2; ++br) {
    const nsac_score_mlp& p = br == 0 ? r : t;
    for (int i = tid; i < HID * NQp; i += nth) {
      const int o = i / NQp, k = i - o * NQp;
      w1[br][i] = __float2half_rn(k < NQ ? p.w1[(size_t)o * NQ + k] : 0.f);
    }
    for (int i = tid; i < HID * HID; i += nth) w2[br][i] = __float2half_rn(p.w2[i]);
    for (int k = tid; k < HID; k += nth) {
      vecs[br * HID + k] = p.b1[k];
      vecs[2 * HID + br * HID + k] = p.b2[k];
      float s = 0.f;
      for (int o = 0; o < 64; ++o) s = fmaf(p.w4[o], p.w3[o * HID + k], s);
      vecs[4 * HID + br * HID + k] = s;
    }
    if (tid == 0) {
      float s = p.b4[0];
      for (int o = 0; o < 64; ++o) s = fmaf(p.w4[o], p.b3[o], s);
      c34[br] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------ selection kernel
__device__ int block_arg_extreme(const float* vals, int n, bool want_max, float* redv, int* redi) {
  float best = want_max ? -INFINITY : INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = vals[i];
    if ((want_max ? v > best : v < best)) { best = v; bi = i; }
  }
  __syncthreads();
  redv[threadIdx.x] = best;
  redi[threadIdx.x] = bi;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int t = 1; t < blockDim.x; ++t) {
      const float v = redv[t];
      const int i = redi[t];
      if (i == 0x7fffffff) continue;
      if ((want_max ? v > best : v < best) || (v == best && i < bi)) { best = v; bi = i; }
    }
    redi[0] = bi == 0x7fffffff ? 0 : bi;      // no finite candidate (all NaN): hypothesis 0, never an out-of-range index
  }
  __syncthreads();
  const int r = redi[0];
  __syncthreads();
  return r;
}

constexpr int SEL_THREADS = 256;

struct SelParams {
  const float* q_h; const float* t_h; const float* q0; const float* t0;
  const float* feat_rot0; const float* feat_tran0; const int32_t* matched_num;
  const float* w_rots; const float* b_rots; const float* w_trans; const float* b_trans;
  float* logits; float* sums; const float* partials;
  int B, NQ, tiles_per_pair, out_cam_type;
  float* pose; float* score_rot; float* score_tran; int32_t* sel_idx;
  // fused result exchange: the 64-byte result row of pair b is also stored into row (row_offset + b) of every
  // rank's [world*B, 16] buffer through NVLink peer mappings (replaces the all-gather collective)
  float* const* peer_rows; int num_peers; int row_offset;
};

__device__ __forceinline__ void publish_row(const SelParams& p, int b, const float* P) {
  __syncthreads();
  if (p.peer_rows && threadIdx.x < 16) {
    const float v = P[threadIdx.x];
    for (int r = 0; r < p.num_peers; ++r) p.peer_rows[r][(size_t)(p.row_offset + b) * 16 + threadIdx.x] = v;
    __threadfence_system();
  }
}

__global__ void __launch_bounds__(SEL_THREADS)
score_select_tc_kernel(const SelParams p) {
  extern __shared__ float sm[];
  const int C = C_FEAT;
  const int b = blockIdx.x, tid = threadIdx.x, H1n = p.NQ + 1;
  float* fe = sm;                      // [4][256]: avg_rot, avg_tran, soft_rot, soft_tran
  float* red = fe + 4 * C;             // [SEL_THREADS]
  int* redi = reinterpret_cast<int*>(red + SEL_THREADS);
  float* outv = reinterpret_cast<float*>(redi + SEL_THREADS);   // [16]
  float* misc = outv + 16;             // [8]: l0[2], M[2], S[2], d0[2]

  const int m = p.matched_num[b];
  float* P = p.pose + (size_t)b * 16;
  pdl_wait();                       // launched early (programmatic dependent launch): the tile kernel's results from here on
  if (p.score_rot) for (int h = tid; h < H1n; h += blockDim.x) p.score_rot[(size_t)b * H1n + h] = 0.f;
  if (p.score_tran) for (int h = tid; h < H1n; h += blockDim.x) p.score_tran[(size_t)b * H1n + h] = 0.f;
  if (p.sel_idx && tid < 2) p.sel_idx[b * 2 + tid] = -1;
  if (m == 0) {  // :964-969
    if (tid < 3) { P[tid] = p.t0[b * 3 + tid]; P[7 + tid] = p.t0[b * 3 + tid]; }
    if (tid < 4) { P[3 + tid] = p.q0[b * 4 + tid]; P[10 + tid] = p.q0[b * 4 + tid]; }
    if (tid == 0) { P[14] = 0.f; P[15] = 0.f; }
    publish_row(p, b, P);
    return;
  }
  // ---- merge the tile partials with hypothesis 0 (log-sum-exp rescale).  The kernel is a chain of dependent global
  // round trips, so every thread issues ALL of its loads right after m is known (headers and hypothesis-0 logits are
  // the same addresses for the whole block: broadcast) and computes M and S redundantly instead of thread 0 + barriers.
  const int ntile = (m + TILE_H - 1) / TILE_H;
  const float* part0 = p.partials + ((size_t)b * p.tiles_per_pair * 2) * PART_STRIDE;
  const float l0r = p.logits[(size_t)b * H1n], l0t = p.logits[(size_t)p.B * H1n + (size_t)b * H1n];   // hypothesis 0 (row-0 tiles)
  const float f0r = p.feat_rot0[(size_t)b * C + tid], f0t = p.feat_tran0[(size_t)b * C + tid];         // tid = channel
  float Mr_ = l0r, Mt_ = l0t;                  // every logit here lacks the constant c34 (softmax-invariant)
  for (int t = 0; t < ntile; ++t) {
    Mr_ = fmaxf(Mr_, part0[(t * 2 + 0) * PART_STRIDE]);
    Mt_ = fmaxf(Mt_, part0[(t * 2 + 1) * PART_STRIDE]);
  }
  float Sr_ = expf(l0r - Mr_), St_ = expf(l0t - Mt_);
  float ar = 0.f, at = 0.f, wr = 0.f, wt = 0.f;      // aggregated features of channel tid (:1047-1087)
  for (int t = 0; t < ntile; ++t) {
    const float* pr = part0 + (t * 2 + 0) * PART_STRIDE;
    const float* pt = pr + PART_STRIDE;
    const float er = expf(pr[0] - Mr_), et = expf(pt[0] - Mt_);
    Sr_ = fmaf(er, pr[1], Sr_);
    St_ = fmaf(et, pt[1], St_);
    wr = fmaf(er, pr[PART_HDR + tid], wr);
    wt = fmaf(et, pt[PART_HDR + tid], wt);
    ar += pr[PART_HDR + C + tid];
    at += pt[PART_HDR + C + tid];
  }
  if (tid == 0) { misc[0] = l0r; misc[1] = l0t; }
  // scores (softmax over hypotheses 0..m, :1010-1014, :1039-1043) -> global (optional) + kept for max-score
  float* lr = p.logits + (size_t)b * H1n;
  float* lt = p.logits + (size_t)p.B * H1n + (size_t)b * H1n;
  // every thread has read the hypothesis-0 logits (l0r / l0t above) before thread 0 overwrites lr[0] / lt[0] with scores
  __syncthreads();
  if (p.score_rot || p.score_tran || (m > 1 && p.out_cam_type == NSAC_CAM_MAX_SCORE)) {
    for (int h = tid; h <= m; h += blockDim.x) {
      const float a = expf(lr[h] - Mr_) / Sr_;
      const float c = expf(lt[h] - Mt_) / St_;
      lr[h] = a; lt[h] = c;                     // logits buffer now holds the scores
      if (p.score_rot) p.score_rot[(size_t)b * H1n + h] = a;
      if (p.score_tran) p.score_tran[(size_t)b * H1n + h] = c;
    }
  }
  __syncthreads();
  int sel_r = -1, sel_t = -1;
  if (m > 1 && p.out_cam_type == NSAC_CAM_MIN_COST) {
    sel_r = block_arg_extreme(p.sums + (size_t)b * H1n, m + 1, false, red, redi);
    sel_t = block_arg_extreme(p.sums + (size_t)p.B * H1n + (size_t)b * H1n, m + 1, false, red, redi);
  } else if (m > 1 && p.out_cam_type == NSAC_CAM_MAX_SCORE) {
    sel_r = block_arg_extreme(lr, m + 1, true, red, redi);
    sel_t = block_arg_extreme(lt, m + 1, true, red, redi);
  }
  if (p.sel_idx && tid == 0) { p.sel_idx[b * 2] = sel_r; p.sel_idx[b * 2 + 1] = sel_t; }
  {
    const int c = tid;
    wr = (wr + expf(l0r - Mr_) * f0r) / Sr_;
    wt = (wt + expf(l0t - Mt_) * f0t) / St_;
    if (m > 1) {  // the initial pose joins the average only when m > 1 (:1052-1063)
      const float w = 1.f / (float)(m + 1);
      ar = (ar + f0r) * w;
      at = (at + f0t) * w;
    }
    fe[c] = ar; fe[C + c] = at; fe[2 * C + c] = wr; fe[3 * C + c] = wt;
  }
  __syncthreads();
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int o = warp; o < 14; o += (blockDim.x >> 5)) {
      const float* f; const float* w; float bias;
      if (o < 4)       { f = fe;         w = p.w_rots + o * C;         bias = p.b_rots[o]; }
      else if (o < 7)  { f = fe + C;     w = p.w_trans + (o - 4) * C;  bias = p.b_trans[o - 4]; }
      else if (o < 11) { f = fe + 2 * C; w = p.w_rots + (o - 7) * C;   bias = p.b_rots[o - 7]; }
      else             { f = fe + 3 * C; w = p.w_trans + (o - 11) * C; bias = p.b_trans[o - 11]; }
      float a = 0.f;
      for (int c = lane; c < C; c += 32) a = fmaf(f[c], w[c], a);
      a = warp_sum(a);
      if (lane == 0) outv[o] = a + bias;
    }
  }
  __syncthreads();
  if (tid == 0) {
    float qa[4] = {outv[0], outv[1], outv[2], outv[3]};
    const float na = fmaxf(sqrtf(qa[0] * qa[0] + qa[1] * qa[1] + qa[2] * qa[2] + qa[3] * qa[3]), 1e-12f);
    for (int i = 0; i < 4; ++i) qa[i] /= na;
    const float ta[3] = {outv[4], outv[5], outv[6]};
    float qf[4], tf[3];
    if (m <= 1 || p.out_cam_type == NSAC_CAM_AVG_ALL) {
      for (int i = 0; i < 4; ++i) qf[i] = qa[i];
      for (int i = 0; i < 3; ++i) tf[i] = ta[i];
    } else if (p.out_cam_type == NSAC_CAM_SOFT) {
      float qs[4] = {outv[7], outv[8], outv[9], outv[10]};
      const float ns = fmaxf(sqrtf(qs[0] * qs[0] + qs[1] * qs[1] + qs[2] * qs[2] + qs[3] * qs[3]), 1e-12f);
      for (int i = 0; i < 4; ++i) qf[i] = qs[i] / ns;
      for (int i = 0; i < 3; ++i) tf[i] = outv[11 + i];
    } else {
      const float* q = sel_r == 0 ? p.q0 + (size_t)b * 4 : p.q_h + ((size_t)b * p.NQ + sel_r - 1) * 4;
      const float* t = sel_t == 0 ? p.t0 + (size_t)b * 3 : p.t_h + ((size_t)b * p.NQ + sel_t - 1) * 3;
      for (int i = 0; i < 4; ++i) qf[i] = q[i];
      for (int i = 0; i < 3; ++i) tf[i] = t[i];
    }
    for (int i = 0; i < 3; ++i) { P[i] = tf[i]; P[7 + i] = ta[i]; }
    for (int i = 0; i < 4; ++i) { P[3 + i] = qf[i]; P[10 + i] = qa[i]; }
    P[14] = (float)m;
    P[15] = 0.f;
  }
  publish_row(p, b, P);
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}
bool make_map_f16(CUtensorMap* map, const void* base, int rows, int cols) {   // row-major fp16 [rows, cols], box 128 x 64
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)KB, (cuuint32_t)TILE_H};
  cuuint32_t es[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
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
inline int nq_padded(int NQ) { return (NQ + KB - 1) / KB * KB; }
inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }
}  // namespace

static unsigned long long* g_score_trace = nullptr;
static int g_score_trace_cta = 0;
// Debug aid: the next nsac_score_aggregate_tc launches record the role timeline of CTA `cta` into `buf`
// (8 * 256 uint64 device words, zero-filled by the caller); nullptr switches it off.
extern "C" int nsac_debug_score_trace(void* buf, int cta) {
  g_score_trace = static_cast<unsigned long long*>(buf);
  g_score_trace_cta = cta;
  return NSAC_OK;
}

extern "C" size_t nsac_score_pack_bytes(int NQ) {
  return NQ < 1 ? 0 : align256(pack_total_bytes(nq_padded(NQ)));
}

extern "C" int nsac_score_pack(const nsac_score_mlp* rot_mlp, const nsac_score_mlp* tran_mlp, int NQ, void* pack, void* stream) {
  NSAC_REQUIRE(rot_mlp && tran_mlp && pack && NQ >= 1, "nsac_score_pack: bad arguments");
  NSAC_REQUIRE((reinterpret_cast<uintptr_t>(pack) & 255) == 0, "nsac_score_pack: pack buffer must be 256-byte aligned");
  const nsac_score_mlp* mm[2] = {rot_mlp, tran_mlp};
  for (int i = 0; i < 2; ++i)
    NSAC_REQUIRE(mm[i]->w1 && mm[i]->b1 && mm[i]->w2 && mm[i]->b2 && mm[i]->w3 && mm[i]->b3 && mm[i]->w4 && mm[i]->b4,
                 "nsac_score_pack: incomplete score MLP weights");
  score_pack_kernel<<<64, 256, 0, static_cast<cudaStream_t>(stream)>>>(*rot_mlp, *tran_mlp, NQ, nq_padded(NQ), static_cast<uint8_t*>(pack));
  NSAC_CHECK_LAUNCH("score_pack_kernel");
  return NSAC_OK;
}

extern "C" size_t nsac_score_tc_workspace_bytes(int B, int NQ) {
  if (B < 0 || NQ < 1) return 0;
  const size_t per = (size_t)B * (NQ + 1);
  const int tiles = (NQ + TILE_H - 1) / TILE_H;
  return align256(4 * per * sizeof(float)) + align256((size_t)B * tiles * 2 * PART_STRIDE * sizeof(float)) +
         align256((size_t)B * nq_padded(NQ) * 64) + align256((size_t)2 * B * nq_padded(NQ) * sizeof(__half)) +
         align256((size_t)B * tiles * TILE_H * 64) + 256;
}

extern "C" size_t nsac_score_pack_vecs_offset(int NQ) { return NQ < 1 ? 0 : pack_off_vecs(nq_padded(NQ)); }

extern "C" int nsac_score_aggregate_tc(const float* geo_local, const float* q_h, const float* t_h, const float* q0,
                                       const float* t0, const float* feat_rot, const float* feat_tran,
                                       const float* feat_rot0, const float* feat_tran0, const int32_t* matched_num,
                                       const void* pack, const float* w_rots, const float* b_rots, const float* w_trans,
                                       const float* b_trans, int B, int NQ, int out_cam_type, float* pose, float* score_rot,
                                       float* score_tran, int32_t* sel_idx, void* workspace,
                                       float* const* peer_rows, int num_peers, int row_offset, void* stream) {
  return nsac_score_aggregate_tc_cv(geo_local, q_h, t_h, q0, t0, feat_rot, feat_tran, feat_rot0, feat_tran0, matched_num, pack, nullptr,
                                    w_rots, b_rots, w_trans, b_trans, B, NQ, out_cam_type, pose, score_rot, score_tran, sel_idx,
                                    workspace, peer_rows, num_peers, row_offset, stream);
}

extern "C" int nsac_score_aggregate_tc_cv(const float* geo_local, const float* q_h, const float* t_h, const float* q0,
                                          const float* t0, const float* feat_rot, const float* feat_tran,
                                          const float* feat_rot0, const float* feat_tran0, const int32_t* matched_num,
                                          const void* pack, const float* vecs_host, const float* w_rots, const float* b_rots,
                                          const float* w_trans, const float* b_trans, int B, int NQ, int out_cam_type, float* pose,
                                          float* score_rot, float* score_tran, int32_t* sel_idx, void* workspace,
                                          float* const* peer_rows, int num_peers, int row_offset, void* stream) {
  NSAC_REQUIRE(!peer_rows || (num_peers >= 1 && row_offset >= 0), "nsac_score_aggregate_tc: bad peer arguments");
  NSAC_REQUIRE(geo_local && q_h && t_h && q0 && t0 && feat_rot && feat_tran && feat_rot0 && feat_tran0 && matched_num &&
                   pack && w_rots && b_rots && w_trans && b_trans && pose && workspace,
               "nsac_score_aggregate_tc: null pointer");
  NSAC_REQUIRE(B >= 0 && NQ >= 1, "nsac_score_aggregate_tc: bad shape B=%d NQ=%d", B, NQ);
  NSAC_REQUIRE(out_cam_type >= 0 && out_cam_type <= 3, "nsac_score_aggregate_tc: bad out_cam_type %d", out_cam_type);
  NSAC_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0 && (reinterpret_cast<uintptr_t>(pack) & 255) == 0,
               "nsac_score_aggregate_tc: workspace / pack must be 256-byte aligned");
  NSAC_REQUIRE((reinterpret_cast<uintptr_t>(feat_rot) & 15) == 0 && (reinterpret_cast<uintptr_t>(feat_tran) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(q_h) & 15) == 0 && (reinterpret_cast<uintptr_t>(q0) & 15) == 0,
               "nsac_score_aggregate_tc: feature / quaternion tensors must be 16-byte aligned");
  if (B == 0) return NSAC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int NQp = nq_padded(NQ), tiles = (NQ + TILE_H - 1) / TILE_H;
  const size_t per = (size_t)B * (NQ + 1);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const uint8_t* pk = static_cast<const uint8_t*>(pack);
  float* logits = reinterpret_cast<float*>(ws);
  float* sums = logits + 2 * per;
  float* partials = reinterpret_cast<float*>(ws + align256(4 * per * sizeof(float)));
  uint8_t* cjk = ws + align256(4 * per * sizeof(float)) + align256((size_t)B * tiles * 2 * PART_STRIDE * sizeof(float));
  __half* x0 = reinterpret_cast<__half*>(cjk + align256((size_t)B * NQp * 64));
  uint4* afr = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(x0) + align256((size_t)2 * B * NQp * sizeof(__half)));
  const float* vecs = reinterpret_cast<const float*>(pk + pack_off_vecs(NQp));
  const int row0_tiles = nsac_cdiv(B, TILE_H);

  // column blocks of every (pair, k-block) + hypothesis 0 of every pair
  score_prep_kernel<<<2 * B, PREP_THREADS, 0, s>>>(geo_local, q_h, t_h, q0, t0, matched_num, B, NQ, NQp, tiles * TILE_H, cjk, x0, afr, sums);
  NSAC_CHECK_LAUNCH("score_prep_kernel");

  CUtensorMap m1r, m1t, m2r, m2t, mx0r, mx0t;
  const bool ok = make_map_f16(&m1r, pk, HID, NQp) && make_map_f16(&m1t, pk + pack_w1_bytes(NQp), HID, NQp) &&
                  make_map_f16(&m2r, pk + 2 * pack_w1_bytes(NQp), HID, HID) &&
                  make_map_f16(&m2t, pk + 2 * pack_w1_bytes(NQp) + HID * HID * 2, HID, HID) &&
                  make_map_f16(&mx0r, x0, B, NQp) && make_map_f16(&mx0t, x0 + (size_t)B * NQp, B, NQp);
  if (!ok) {
    nsac_set_error("nsac_score_aggregate_tc: cuTensorMapEncodeTiled failed");
    return NSAC_ERR_LAUNCH;
  }
  TcParams tp;
  tp.trace = g_score_trace; tp.trace_cta = g_score_trace_cta;
  tp.geo_local = geo_local; tp.q_h = q_h; tp.t_h = t_h; tp.q0 = q0; tp.t0 = t0; tp.feat_rot = feat_rot; tp.feat_tran = feat_tran;
  tp.matched_num = matched_num; tp.vecs = vecs; tp.cjk = cjk; tp.afr = afr;
  tp.B = B; tp.NQ = NQ; tp.NQp = NQp; tp.tiles_per_pair = tiles;
  tp.need_sums = out_cam_type == NSAC_CAM_MIN_COST; tp.logits = logits; tp.sums = sums; tp.partials = partials;
  static bool attr = false;
  if (!attr) {
    NSAC_CUDA(cudaFuncSetAttribute(score_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    NSAC_CUDA(cudaFuncSetAttribute(score_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    NSAC_CUDA(cudaFuncSetAttribute(score_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    NSAC_CUDA(cudaFuncSetAttribute(score_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr = true;
  }
  if (vecs_host) memcpy(tp.cvec, vecs_host, sizeof(tp.cvec));      // host mirror of the pack's vectors -> kernel parameters
  // Item list: B*tiles hypothesis tiles + the row-0 tiles.  Items go round-robin to the persistent CTAs, so the
  // (slightly more expensive) row-0 tiles are inserted where the CTAs that get one item fewer pick them up.
  const int items = B * tiles + row0_tiles;
  const int grid = items < sm_count() ? items : sm_count();
  const int rem = items % grid;
  tp.num_items = items; tp.row0_tiles = row0_tiles;
  tp.row0_at = (rem != 0 && rem + row0_tiles <= grid) ? rem : 0;
  // prep -> tiles -> selection are chained by programmatic dependent launch: each kernel may start while its predecessor
  // drains and blocks in griddepcontrol.wait before it touches the predecessor's results (captured as programmatic edges
  // in a CUDA graph)
  cudaLaunchAttribute pdl_attr[1];
  pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl_attr[0].val.programmaticStreamSerializationAllowed = NSAC_SCORE_PDL;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = s;
  cfg.attrs = pdl_attr; cfg.numAttrs = 1;
  if (tp.need_sums && vecs_host)
    NSAC_CUDA(cudaLaunchKernelEx(&cfg, score_tc_kernel<true, true>, m1r, m1t, m2r, m2t, mx0r, mx0t, tp));
  else if (tp.need_sums)
    NSAC_CUDA(cudaLaunchKernelEx(&cfg, score_tc_kernel<true, false>, m1r, m1t, m2r, m2t, mx0r, mx0t, tp));
  else if (vecs_host)
    NSAC_CUDA(cudaLaunchKernelEx(&cfg, score_tc_kernel<false, true>, m1r, m1t, m2r, m2t, mx0r, mx0t, tp));
  else
    NSAC_CUDA(cudaLaunchKernelEx(&cfg, score_tc_kernel<false, false>, m1r, m1t, m2r, m2t, mx0r, mx0t, tp));
  NSAC_CHECK_LAUNCH("score_tc_kernel");

  SelParams sp;
  sp.q_h = q_h; sp.t_h = t_h; sp.q0 = q0; sp.t0 = t0; sp.feat_rot0 = feat_rot0; sp.feat_tran0 = feat_tran0;
  sp.matched_num = matched_num; sp.w_rots = w_rots; sp.b_rots = b_rots; sp.w_trans = w_trans; sp.b_trans = b_trans;
  sp.logits = logits; sp.sums = sums; sp.partials = partials;
  sp.B = B; sp.NQ = NQ; sp.tiles_per_pair = tiles; sp.out_cam_type = out_cam_type; sp.pose = pose; sp.score_rot = score_rot;
  sp.score_tran = score_tran; sp.sel_idx = sel_idx;
  sp.peer_rows = peer_rows; sp.num_peers = num_peers; sp.row_offset = row_offset;
  const size_t sel_smem = sizeof(float) * (4 * C_FEAT + SEL_THREADS + 16 + 8) + sizeof(int) * SEL_THREADS;
  cfg.gridDim = dim3(B); cfg.blockDim = dim3(SEL_THREADS); cfg.dynamicSmemBytes = sel_smem;
  NSAC_CUDA(cudaLaunchKernelEx(&cfg, score_select_tc_kernel, sp));
  NSAC_CHECK_LAUNCH("score_select_tc_kernel");
  return NSAC_OK;
}
