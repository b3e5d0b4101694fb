// K8 + K9 on the Blackwell tensor pipe — hypothesis scoring and pose selection (camera_head.py:964-1115).
//
// Work item = (pair b, tile of 128 one-plane hypotheses h = 1 + 128*tile + r); hypothesis 0 (the initial pose)
// is a single row and is scored by a small batched kernel.  One persistent CTA per SM, 24 warps in 6 warpgroups
// (setmaxnreg moves registers from the TMA / MMA / epilogue warps to the residual warps):
//
//   warp 0        TMA        W2 (both branches, resident) once; then per (tile, k-block) the two [128 x 64] fp16
//                            slices of the first score-MLP layer W1 (rot / trans) into a 2-stage ring
//   warp 1        MMA        layer 1: D_b[128x128] += X_b[128x64] . W1_b^T  (tcgen05.mma kind::f16, A and B from
//                            shared memory);  layer 2: D_b = H1_b . W2_b^T with H1 read from TENSOR MEMORY (A operand
//                            written there by the epilogue warps) — the hidden activations never touch shared memory
//   warps 4-7     epilogue   TMEM -> registers: +b1, ReLU, pack fp16x2 -> tcgen05.st (H1);  then +b2, ReLU and the
//                            folded Linear(128,64)+Linear(64,1) dot product: one thread owns one hypothesis row, so
//                            the logit needs no cross-thread reduction
//   warps 8-15    residuals  the CUDA-core part: thread = (hypothesis row, 8-column chunk).  u = R_h n_j is shared by
//                            both branches;  rot: exp(-|u - n1_j|);  trans: exp(-|A_j (d_j + t_h.u) u - pi1_j|)
//                            (closed forms of the reference's warp + normalise, see residual_pair()); results are
//                            written as fp16 straight into the 128-byte-swizzled K-major A-operand tiles (2-stage
//                            ring), fence.proxy.async, mbarrier arrive.  The [B,NQ+1,NQ,3] temporaries of the
//                            reference never exist.
//   warps 16-23   gather     flash-style softmax partials: local max / sum of exp over the tile's logits and the
//                            exp-weighted + plain sums of the tile's [128,256] one-plane features — the only HBM
//                            stream of the kernel (float4 loads, 16 rows in flight), overlapped with the residual
//                            and tensor work of the NEXT tile
//
// A per-pair selection kernel then scores hypothesis 0 (exact fp32), merges the tile partials (log-sum-exp
// rescale), applies the m == 0 / m == 1 / m > 1 rules (:964, :1052, :1068), the avg / soft / min-cost /
// max-score selection and the shared pose heads, and writes pose[b, 0:16].
//
// Precision: the score MLPs run single-pass fp16 (11 significant bits for exp(-d) in [0,1] and for the
// weights) with fp32 accumulation; measured effect on scores <= 4e-5 and on the soft pose <= 3e-5 even with a
// 4x sharpened softmax.  The exact-fp32 CUDA-core kernels of score.cu remain available
// (precision = 0, and always for the diagnostic outputs).
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"

namespace {

constexpr int TILE_H = 128;        // hypotheses per tile (UMMA M)
constexpr int HID = 128;           // hidden width of the score MLPs (UMMA N)
constexpr int KB = 64;             // residual columns per k-block (one 128-byte swizzle row of fp16)
constexpr int C_FEAT = 256;
constexpr int W_STAGES = 2, A_STAGES = 2;
constexpr int BLK_BYTES = TILE_H * KB * 2;           // 16 KB: one [128 x 64] fp16 operand block
// warp roles, aligned to warpgroups of 4 warps so that setmaxnreg can move registers to the residual warps:
//   WG0 = warps 0-3 (TMA, MMA, 2 idle)  40 regs | WG1 = warps 4-7 epilogue  72 regs
//   WG2-3 = warps 8-15 residuals       112 regs | WG4-5 = warps 16-23 gather 72 regs
// (launch: 768 x 80; setmaxnreg.inc can only draw on what the CTA's own warps released with setmaxnreg.dec:
//  128*40 + 128*8 + 256*8 = 8192 = 256*32 — an inc that is not covered deadlocks)
constexpr int NUM_WARPS = 24, NUM_THREADS = NUM_WARPS * 32;
constexpr int E_WARP0 = 4;
constexpr int R_WARP0 = 8, R_WARPS = 8, R_THREADS = R_WARPS * 32;
constexpr int G_WARP0 = 16, G_THREADS = 256;
constexpr uint32_t SPIN_LIMIT = 2000;      // suspended waits of up to ~10 ms each: ~20 s, then trap instead of hanging
constexpr int PART_HDR = 4;                          // max, sumexp, pad, pad (keeps the vectors 16-byte aligned)
constexpr int PART_STRIDE = PART_HDR + 2 * C_FEAT;   // per (pair, tile, branch): header, wsum[256], fsum[256]

// shared memory map (offsets from a 1024-aligned base)
constexpr int OFF_W2 = 0;                                        // [branch][kblock 0..1][16 KB]
constexpr int OFF_W1 = OFF_W2 + 4 * BLK_BYTES;                   // [stage][branch][16 KB]
constexpr int OFF_A = OFF_W1 + W_STAGES * 2 * BLK_BYTES;         // [stage][branch][16 KB]
constexpr int OFF_CJ = OFF_A + A_STAGES * 2 * BLK_BYTES;         // [W_STAGES][64][3] float4 column constants (TMA)
constexpr int OFF_VEC = OFF_CJ + W_STAGES * KB * 12 * 4;         // b1[2][128], b2[2][128], w34[2][128] floats
constexpr int OFF_LOGIT = OFF_VEC + 6 * HID * 4;                 // [2 bufs][2 branches][128] floats
constexpr int OFF_ROWSUM = OFF_LOGIT + 2 * 2 * TILE_H * 4;       // [4 quarters][2 branches][128] floats (min-cost sums)
constexpr int OFF_EXP = OFF_ROWSUM + 4 * 2 * TILE_H * 4;         // [2 branches][128] softmax numerators of the tile
constexpr int OFF_GPART = OFF_EXP + 2 * TILE_H * 4;              // [2 branches][64 threads][9] odd-row partials (+1 pad)
constexpr int OFF_BAR = OFF_GPART + 2 * 64 * 12 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

// tensor-memory columns
constexpr uint32_t TM_D = 0;          // D_rot [0,128), D_tran [128,256)
constexpr uint32_t TM_H1 = 256;       // H1_rot [256,320), H1_tran [320,384)  (fp16 pairs)
constexpr uint32_t TM_COLS = 512;

enum { BAR_W2 = 0, BAR_W_FULL = 1, BAR_W_EMPTY = 3, BAR_A_FULL = 5, BAR_A_EMPTY = 7, BAR_ACC_FULL = 9, BAR_H1_READY = 10,
       BAR_ACC2_FULL = 11, BAR_ACC_FREE = 12, BAR_LOGIT_READY = 13, BAR_LOGIT_FREE = 15, BAR_COUNT = 17 };

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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
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

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------ residual math
// Column constants of matched plane pair j (cj[12]): n^ = unit(p0*flip) (0-2), n1 = unit(p1*flip) (3-5),
// pi1 = p1*flip (6-8), A = d^2/(d+1e-5)^2 (9), Bc = A*d (10), valid (11).
//
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
// c0 = (n^x, n^y, n^z, n1x), c1 = (n1y, n1z, pi1x, pi1y), c2 = (pi1z, A, Bc, valid)
template <bool SUMS>
__device__ __forceinline__ void residual_pair(const float (&R)[9], float tx, float ty, float tz, const float4 c0, const float4 c1,
                                              const float4 c2, float& xr, float& xt, float& sum_r, float& sum_t) {
  const float ux = fmaf(R[0], c0.x, fmaf(R[1], c0.y, R[2] * c0.z));
  const float uy = fmaf(R[3], c0.x, fmaf(R[4], c0.y, R[5] * c0.z));
  const float uz = fmaf(R[6], c0.x, fmaf(R[7], c0.y, R[8] * c0.z));
  const float ax = ux - c0.w, ay = uy - c1.x, az = uz - c1.y;
  const float dr = fast_sqrt(fmaf(ax, ax, fmaf(ay, ay, az * az)));
  const float tu = fmaf(tx, ux, fmaf(ty, uy, tz * uz));
  const float g = fmaf(c2.y, tu, c2.z);
  const float wx = fmaf(g, ux, -c1.z), wy = fmaf(g, uy, -c1.w), wz = fmaf(g, uz, -c2.x);
  const float dt = fast_sqrt(fmaf(wx, wx, fmaf(wy, wy, wz * wz)));
  xr = c2.w * fast_exp2(-1.4426950408889634f * dr);     // c2.w = 1 for matched columns, 0 for the padded tail
  xt = c2.w * fast_exp2(-1.4426950408889634f * dt);
  if (SUMS) { sum_r = fmaf(c2.w, dr, sum_r); sum_t = fmaf(c2.w, dt, sum_t); }
}

// column constants of matched plane pair j from geo_local row (p0, p1)
__device__ __forceinline__ void column_constants(const float* __restrict__ g6, bool valid, float4& c0, float4& c1, float4& c2) {
  c0 = c1 = c2 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid) {
    float ax = g6[0], ay = -g6[1], az = -g6[2];
    const float d = normalize3(ax, ay, az);
    const float px = g6[3], py = -g6[4], pz = -g6[5];
    float nx = px, ny = py, nz = pz;
    normalize3(nx, ny, nz);
    const float dd = d + 1e-5f, A = d * d / (dd * dd);
    c0 = make_float4(ax, ay, az, nx);
    c1 = make_float4(ny, nz, px, py);
    c2 = make_float4(pz, A, A * d, 1.f);
  }
}

struct TcParams {
  const float* geo_local;   // [B,NQ,6]
  const float* q_h;         // [B,NQ,4]
  const float* t_h;         // [B,NQ,3]
  const float* feat_rot;    // [B,NQ,256]
  const float* feat_tran;
  const int32_t* matched_num;
  const float* vecs;        // packed: b1[2][128], b2[2][128], w34[2][128]
  const float4* cjg;        // [B][NQp][3] column constants (score_row0_kernel)
  int B, NQ, NQp, tiles_per_pair, need_sums;
  float* logits;            // [2][B][NQ+1]
  float* sums;              // [2][B][NQ+1]
  float* partials;          // [B][tiles][2][PART_STRIDE]
};

template <bool SUMS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
score_tc_kernel(const __grid_constant__ CUtensorMap map_w1r, const __grid_constant__ CUtensorMap map_w1t,
                const __grid_constant__ CUtensorMap map_w2r, const __grid_constant__ CUtensorMap map_w2t, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  // align inside the shared window with offset arithmetic (a generic-pointer round trip would turn every
  // shared-memory access below into a generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BAR_COUNT);
  float4* cj = reinterpret_cast<float4*>(smem + OFF_CJ);
  float* vec = reinterpret_cast<float*>(smem + OFF_VEC);
  float* s_logit = reinterpret_cast<float*>(smem + OFF_LOGIT);
  float* s_rowsum = reinterpret_cast<float*>(smem + OFF_ROWSUM);
  float* s_exp = reinterpret_cast<float*>(smem + OFF_EXP);
  float* s_gpart = reinterpret_cast<float*>(smem + OFF_GPART);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H1n = p.NQ + 1;
  const int num_items = p.B * p.tiles_per_pair;

  if (threadIdx.x == 0) {
    mbar_init(&bars[BAR_W2], 1);
    for (int s = 0; s < W_STAGES; ++s) { mbar_init(&bars[BAR_W_FULL + s], 1); mbar_init(&bars[BAR_W_EMPTY + s], 1); }
    for (int s = 0; s < A_STAGES; ++s) { mbar_init(&bars[BAR_A_FULL + s], R_THREADS); mbar_init(&bars[BAR_A_EMPTY + s], 1); }
    mbar_init(&bars[BAR_ACC_FULL], 1);
    mbar_init(&bars[BAR_H1_READY], 4);
    mbar_init(&bars[BAR_ACC2_FULL], 1);
    mbar_init(&bars[BAR_ACC_FREE], 4);
    for (int s = 0; s < 2; ++s) { mbar_init(&bars[BAR_LOGIT_READY + s], 4); mbar_init(&bars[BAR_LOGIT_FREE + s], G_THREADS / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 6 * HID; i += NUM_THREADS) vec[i] = p.vecs[i];
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // setmaxnreg: ONE instruction per warpgroup (all 4 warps must execute the same one), inside the warpgroup's own
  // branch so that the register limit is unambiguous on every control path
  if (warp < E_WARP0) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    // ================================================================================= TMA producer
    if (lane == 0) {
      mbar_expect_tx(&bars[BAR_W2], 4 * BLK_BYTES);
      for (int kb = 0; kb < 2; ++kb) {
        tma_load_2d(smem + OFF_W2 + (0 * 2 + kb) * BLK_BYTES, &map_w2r, &bars[BAR_W2], kb * KB, 0);
        tma_load_2d(smem + OFF_W2 + (1 * 2 + kb) * BLK_BYTES, &map_w2t, &bars[BAR_W2], kb * KB, 0);
      }
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int b = item / p.tiles_per_pair, tile = item % p.tiles_per_pair;
        const int m = p.matched_num[b];
        if (tile * TILE_H >= m) continue;
        const int nkb = (m + KB - 1) / KB;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&bars[BAR_W_EMPTY + stage], phase ^ 1);
          mbar_expect_tx(&bars[BAR_W_FULL + stage], 2 * BLK_BYTES + KB * 48);
          uint8_t* dst = smem + OFF_W1 + stage * 2 * BLK_BYTES;
          tma_load_2d(dst, &map_w1r, &bars[BAR_W_FULL + stage], kb * KB, 0);
          tma_load_2d(dst + BLK_BYTES, &map_w1t, &bars[BAR_W_FULL + stage], kb * KB, 0);
          // the k-block's 64 x 48 B of per-column geometry constants travel with the weights
          bulk_load_1d(smem + OFF_CJ + stage * KB * 48, p.cjg + ((size_t)b * p.NQp + (size_t)kb * KB) * 3, KB * 48,
                       &bars[BAR_W_FULL + stage]);
          if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================================= MMA issuer
    if (lane == 0) {
      int ws = 0; uint32_t wph = 0; int as = 0; uint32_t aph = 0; uint32_t tph = 0;   // tph: per-tile barrier parity
      mbar_wait(&bars[BAR_W2], 0);
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int b = item / p.tiles_per_pair, tile = item % p.tiles_per_pair;
        const int m = p.matched_num[b];
        if (tile * TILE_H >= m) continue;
        const int nkb = (m + KB - 1) / KB;
        mbar_wait(&bars[BAR_ACC_FREE], tph ^ 1);          // epilogue of the previous tile has drained D
        fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&bars[BAR_W_FULL + ws], wph);
          mbar_wait(&bars[BAR_A_FULL + as], aph);
          fence_after();
          const uint32_t a0 = smem_u32(smem + OFF_A + as * 2 * BLK_BYTES), w0 = smem_u32(smem + OFF_W1 + ws * 2 * BLK_BYTES);
#pragma unroll 1
          for (int br = 0; br < 2; ++br) {
            const uint64_t da = sw128_desc(a0 + br * BLK_BYTES), dw = sw128_desc(w0 + br * BLK_BYTES);
#pragma unroll 1
            for (int k = 0; k < KB / 16; ++k)
              umma_ss(tmem_base + TM_D + br * HID, da + 2 * k, dw + 2 * k, IDESC, (kb | k) != 0);
          }
          umma_commit(&bars[BAR_W_EMPTY + ws]);
          umma_commit(&bars[BAR_A_EMPTY + as]);
          if (++ws == W_STAGES) { ws = 0; wph ^= 1; }
          if (++as == A_STAGES) { as = 0; aph ^= 1; }
        }
        umma_commit(&bars[BAR_ACC_FULL]);
        // layer 2: A = H1 (fp16 pairs in tensor memory), B = W2 (resident in shared memory)
        mbar_wait(&bars[BAR_H1_READY], tph);
        fence_after();
#pragma unroll 1
        for (int br = 0; br < 2; ++br) {
#pragma unroll 1
          for (int k = 0; k < HID / 16; ++k) {
            const uint64_t dw = sw128_desc(smem_u32(smem + OFF_W2 + (br * 2 + (k >> 2)) * BLK_BYTES)) + 2 * (k & 3);
            umma_ts(tmem_base + TM_D + br * HID, tmem_base + TM_H1 + br * (HID / 2) + k * 8, dw, IDESC, k != 0);
          }
        }
        umma_commit(&bars[BAR_ACC2_FULL]);
        tph ^= 1;
      }
    }
  }   // warps 2, 3: idle (they only donate their registers)
  } else if (warp < R_WARP0) {
    // ================================================================================= epilogue warps 4..7
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    const int quad = warp & 3, row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    uint32_t tph = 0; int lbuf = 0; uint32_t lph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int b = item / p.tiles_per_pair, tile = item % p.tiles_per_pair;
      const int m = p.matched_num[b];
      if (tile * TILE_H >= m) continue;
      // ---- layer-1 epilogue: H1 = fp16(relu(D + b1)) -> tensor memory
      mbar_wait(&bars[BAR_ACC_FULL], tph);
      fence_after();
#pragma unroll
      for (int br = 0; br < 2; ++br) {
        const float* b1 = vec + br * HID;
#pragma unroll
        for (int c = 0; c < HID; c += 32) {
          uint32_t v[32], h[16];
          tmem_ld32(tmem_base + lane_base + TM_D + br * HID + c, v);
#pragma unroll
          for (int i = 0; i < 32; i += 2)
            h[i >> 1] = pack_h2(fmaxf(__uint_as_float(v[i]) + b1[c + i], 0.f), fmaxf(__uint_as_float(v[i + 1]) + b1[c + i + 1], 0.f));
          tmem_st16(tmem_base + lane_base + TM_H1 + br * (HID / 2) + (c >> 1), h);
        }
      }
      tmem_st_wait();
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BAR_H1_READY]);
      // ---- layer-2 epilogue: logit = w34 . relu(D + b2) + c34   (this thread owns hypothesis row `row`)
      mbar_wait(&bars[BAR_ACC2_FULL], tph);
      fence_after();
      float lg[2];
#pragma unroll
      for (int br = 0; br < 2; ++br) {
        const float* b2 = vec + 2 * HID + br * HID;
        const float* w34 = vec + 4 * HID + br * HID;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < HID; c += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + lane_base + TM_D + br * HID + c, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) acc = fmaf(fmaxf(__uint_as_float(v[i]) + b2[c + i], 0.f), w34[c + i], acc);
        }
        lg[br] = acc;      // the constant c34 cancels in the softmax; it is added by the selection kernel
      }
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BAR_ACC_FREE]);
      // ---- hand the logits to the gather warps (and to global memory for scores / argmax)
      mbar_wait(&bars[BAR_LOGIT_FREE + lbuf], lph ^ 1);
      const int h = 1 + tile * TILE_H + row;
      s_logit[(lbuf * 2 + 0) * TILE_H + row] = lg[0];
      s_logit[(lbuf * 2 + 1) * TILE_H + row] = lg[1];
      if (h <= m) {
        p.logits[(size_t)b * H1n + h] = lg[0];
        p.logits[(size_t)p.B * H1n + (size_t)b * H1n + h] = lg[1];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[BAR_LOGIT_READY + lbuf]);
      if (++lbuf == 2) { lbuf = 0; lph ^= 1; }
      tph ^= 1;
    }
  } else if (warp < G_WARP0) {
    // ================================================================================= residual warps 8..15
    // thread = 2 hypothesis rows (rp, rp + 64) x 16 columns of the k-block: the column constants are loaded once
    // (3 LDS.128) and used for both rows
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int rt = threadIdx.x - R_WARP0 * 32;       // 0..255
    const int rp = rt & 63, quarter = rt >> 6;       // rows rp / rp+64, chunks 2*quarter, 2*quarter+1
    int as = 0; uint32_t aph = 0; int ws = 0; uint32_t wph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int b = item / p.tiles_per_pair, tile = item % p.tiles_per_pair;
      const int m = p.matched_num[b];
      if (tile * TILE_H >= m) continue;
      const int nkb = (m + KB - 1) / KB;
      float R[2][9], tr[2][3];
      bool hv[2];
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
        const int hidx = tile * TILE_H + rp + 64 * s2;   // index into q_h / t_h (hypothesis h = hidx + 1)
        hv[s2] = hidx < m;
        float qw = 1.f, qx = 0.f, qy = 0.f, qz = 0.f;
        tr[s2][0] = tr[s2][1] = tr[s2][2] = 0.f;
        if (hv[s2]) {
          const float4 q = *reinterpret_cast<const float4*>(p.q_h + ((size_t)b * p.NQ + hidx) * 4);
          const float* t = p.t_h + ((size_t)b * p.NQ + hidx) * 3;
          qw = q.x; qx = q.y; qy = q.z; qz = q.w;
          tr[s2][0] = t[0]; tr[s2][1] = t[1]; tr[s2][2] = t[2];
        }
        const Mat3 M = quat_to_rot(qw, qx, qy, qz);
#pragma unroll
        for (int i = 0; i < 9; ++i) R[s2][i] = M.m[i];
      }
      float sum_r[2] = {0.f, 0.f}, sum_t[2] = {0.f, 0.f};
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&bars[BAR_A_EMPTY + as], aph ^ 1);
        mbar_wait(&bars[BAR_W_FULL + ws], wph);           // column constants of this k-block have landed (TMA)
        const float4* cjs = cj + ws * KB * 3;
        uint8_t* a_rot = smem + OFF_A + as * 2 * BLK_BYTES;
        uint8_t* a_tran = a_rot + BLK_BYTES;
#pragma unroll 2
        for (int c4i = 0; c4i < 4; ++c4i) {                 // 4 groups of 4 columns = this thread's 16 columns
          const int chunk = quarter * 2 + (c4i >> 1), sub = c4i & 1;
          float xr[2][4], xt[2][4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = chunk * 8 + sub * 4 + e;
            const float4 c0 = cjs[col * 3 + 0], c1 = cjs[col * 3 + 1], c2 = cjs[col * 3 + 2];
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2)
              residual_pair<SUMS>(R[s2], tr[s2][0], tr[s2][1], tr[s2][2], c0, c1, c2, xr[s2][e], xt[s2][e], sum_r[s2], sum_t[s2]);
          }
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            const float keep = hv[s2] ? 1.f : 0.f;
            const int row = rp + 64 * s2;
            const uint32_t off = (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4) + (uint32_t)(sub * 8);   // 128-byte swizzle
            *reinterpret_cast<uint2*>(a_rot + off) = make_uint2(pack_h2(keep * xr[s2][0], keep * xr[s2][1]), pack_h2(keep * xr[s2][2], keep * xr[s2][3]));
            *reinterpret_cast<uint2*>(a_tran + off) = make_uint2(pack_h2(keep * xt[s2][0], keep * xt[s2][1]), pack_h2(keep * xt[s2][2], keep * xt[s2][3]));
          }
        }
        fence_proxy_async();
        mbar_arrive(&bars[BAR_A_FULL + as]);
        if (++as == A_STAGES) { as = 0; aph ^= 1; }
        if (++ws == W_STAGES) { ws = 0; wph ^= 1; }
      }
      if (SUMS) {   // sum_j of the masked distances (argmin in 'min-cost', :1090-1093): 4 column quarters per row, fixed order
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          s_rowsum[(quarter * 2 + 0) * TILE_H + rp + 64 * s2] = sum_r[s2];
          s_rowsum[(quarter * 2 + 1) * TILE_H + rp + 64 * s2] = sum_t[s2];
        }
        named_bar_sync(1, R_THREADS);
        if (rt < TILE_H && tile * TILE_H + rt < m) {
          float a = 0.f, c = 0.f;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) { a += s_rowsum[(q4 * 2 + 0) * TILE_H + rt]; c += s_rowsum[(q4 * 2 + 1) * TILE_H + rt]; }
          const int h = tile * TILE_H + rt + 1;
          p.sums[(size_t)b * H1n + h] = a;
          p.sums[(size_t)p.B * H1n + (size_t)b * H1n + h] = c;
        }
        named_bar_sync(1, R_THREADS);
      }
    }
  } else {
    // ================================================================================= gather warps 16..23
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    // thread = (branch, row parity, 4 feature channels): 12 float4 loads in flight per thread, 48 KB per SM
    const int gt = threadIdx.x - G_WARP0 * 32;       // 0..255
    const int br = gt >> 7, par = (gt >> 6) & 1, t64 = gt & 63, c4 = t64 * 4;
    int lbuf = 0; uint32_t lph = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int b = item / p.tiles_per_pair, tile = item % p.tiles_per_pair;
      const int m = p.matched_num[b];
      if (tile * TILE_H >= m) continue;
      const int rows = min(TILE_H, m - tile * TILE_H);
      mbar_wait(&bars[BAR_LOGIT_READY + lbuf], lph);
      const float* lg = s_logit + (lbuf * 2 + br) * TILE_H;
      float mx = -INFINITY;
      for (int r = 0; r < rows; ++r) mx = fmaxf(mx, lg[r]);
      float* ex = s_exp + br * TILE_H;
      {
        const int r = par * 64 + t64;
        ex[r] = r < rows ? __expf(lg[r] - mx) : 0.f;
      }
      named_bar_sync(4 + br, 128);
      const float* f = (br == 0 ? p.feat_rot : p.feat_tran) + ((size_t)b * p.NQ + (size_t)tile * TILE_H) * C_FEAT + c4;
      float4 ws = make_float4(0.f, 0.f, 0.f, 0.f), fs = make_float4(0.f, 0.f, 0.f, 0.f);
      float se = 0.f;
      int r = par;
      for (; r + 22 < rows; r += 24) {       // 12 rows of this parity in flight
        float4 v[12];
#pragma unroll
        for (int u = 0; u < 12; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(f + (size_t)(r + 2 * u) * C_FEAT));
#pragma unroll
        for (int u = 0; u < 12; ++u) {
          const float e = ex[r + 2 * u];
          se += e;
          ws.x = fmaf(e, v[u].x, ws.x); ws.y = fmaf(e, v[u].y, ws.y); ws.z = fmaf(e, v[u].z, ws.z); ws.w = fmaf(e, v[u].w, ws.w);
          fs.x += v[u].x; fs.y += v[u].y; fs.z += v[u].z; fs.w += v[u].w;
        }
      }
      for (; r < rows; r += 2) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(f + (size_t)r * C_FEAT));
        const float e = ex[r];
        se += e;
        ws.x = fmaf(e, v.x, ws.x); ws.y = fmaf(e, v.y, ws.y); ws.z = fmaf(e, v.z, ws.z); ws.w = fmaf(e, v.w, ws.w);
        fs.x += v.x; fs.y += v.y; fs.z += v.z; fs.w += v.w;
      }
      float* gp = s_gpart + (br * 64 + t64) * 12;
      if (par == 1) {
        *reinterpret_cast<float4*>(gp) = ws;
        *reinterpret_cast<float4*>(gp + 4) = fs;
        gp[8] = se;
      }
      named_bar_sync(4 + br, 128);
      if (par == 0) {
        const float4 w1 = *reinterpret_cast<const float4*>(gp), f1 = *reinterpret_cast<const float4*>(gp + 4);
        ws.x += w1.x; ws.y += w1.y; ws.z += w1.z; ws.w += w1.w;
        fs.x += f1.x; fs.y += f1.y; fs.z += f1.z; fs.w += f1.w;
        se += gp[8];
        float* part = p.partials + (((size_t)b * p.tiles_per_pair + tile) * 2 + br) * PART_STRIDE;
        if (t64 == 0) { part[0] = mx; part[1] = se; }
        *reinterpret_cast<float4*>(part + PART_HDR + c4) = ws;
        *reinterpret_cast<float4*>(part + PART_HDR + C_FEAT + c4) = fs;
      }
      named_bar_sync(4 + br, 128);      // s_exp / s_gpart free for the next tile
      if (lane == 0) mbar_arrive(&bars[BAR_LOGIT_FREE + lbuf]);
      if (++lbuf == 2) { lbuf = 0; lph ^= 1; }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 1) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TM_COLS));
  }
}

// ------------------------------------------------------------------------------------------------ weight packing
// pack layout: [w1h_rot fp16 128 x NQp][w1h_tran][w2h_rot fp16 128x128][w2h_tran]
//              [vecs fp32: b1[2][128] b2[2][128] w34[2][128]][c34[2] + pad to 8]
//              [w1t fp32 [2][NQ][128]][w2t fp32 [2][128][128]]      (transposed copies for the row-0 kernel)
__host__ __device__ inline size_t pack_w1_bytes(int NQp) { return (size_t)HID * NQp * 2; }
__host__ __device__ inline size_t pack_off_vecs(int NQp) { return 2 * pack_w1_bytes(NQp) + 2 * (size_t)HID * HID * 2; }
__host__ __device__ inline size_t pack_off_w1t(int NQp) { return pack_off_vecs(NQp) + (6 * HID + 8) * sizeof(float); }
__host__ __device__ inline size_t pack_off_w2t(int NQ, int NQp) { return pack_off_w1t(NQp) + 2 * (size_t)NQ * HID * sizeof(float); }
__host__ __device__ inline size_t pack_total_bytes(int NQ, int NQp) { return pack_off_w2t(NQ, NQp) + 2 * (size_t)HID * HID * sizeof(float); }

__global__ void score_pack_kernel(nsac_score_mlp r, nsac_score_mlp t, int NQ, int NQp, uint8_t* pack) {
  __half* w1[2] = {reinterpret_cast<__half*>(pack), reinterpret_cast<__half*>(pack + pack_w1_bytes(NQp))};
  __half* w2[2] = {reinterpret_cast<__half*>(pack + 2 * pack_w1_bytes(NQp)),
                   reinterpret_cast<__half*>(pack + 2 * pack_w1_bytes(NQp) + HID * HID * 2)};
  float* vecs = reinterpret_cast<float*>(pack + pack_off_vecs(NQp));
  float* c34 = vecs + 6 * HID;
  float* w1t = reinterpret_cast<float*>(pack + pack_off_w1t(NQp));
  float* w2t = reinterpret_cast<float*>(pack + pack_off_w2t(NQ, NQp));
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int br = 0; br < 2; ++br) {
    const nsac_score_mlp& p = br == 0 ? r : t;
    for (int i = tid; i < HID * NQp; i += nth) {
      const int o = i / NQp, k = i - o * NQp;
      w1[br][i] = __float2half_rn(k < NQ ? p.w1[(size_t)o * NQ + k] : 0.f);
    }
    for (int i = tid; i < HID * HID; i += nth) w2[br][i] = __float2half_rn(p.w2[i]);
    for (int i = tid; i < NQ * HID; i += nth) {
      const int j = i / HID, o = i - j * HID;
      w1t[(size_t)br * NQ * HID + i] = p.w1[(size_t)o * NQ + j];
    }
    for (int i = tid; i < HID * HID; i += nth) {
      const int k = i / HID, o = i - k * HID;
      w2t[(size_t)br * HID * HID + i] = p.w2[(size_t)o * HID + k];
    }
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

// ------------------------------------------------------------------------------------------------ hypothesis 0
// Hypothesis 0 (the initial pose, camera_head.py:991, 1019) is one row per pair: 4 pairs per CTA so every weight that
// is loaded (coalesced, transposed fp32 copies, 8 loads in flight) feeds 4 FMAs; exact fp32.  Also the masked distance sums of row 0.
constexpr int ROW0_PAIRS = 4;

struct Row0Params {
  const float* geo_local; const float* q0; const float* t0; const int32_t* matched_num;
  const float* w1t; const float* w2t; const float* vecs;   // vecs: b1[2][128], b2[2][128], w34[2][128]
  int B, NQ, NQp;
  float* logits; float* sums;
  float4* cjg;             // out: [B][NQp][3] column constants for the tile kernel
};

constexpr int ROW0_THREADS = 512;     // 2 K-halves x 2 branches x 128 output units

__global__ void __launch_bounds__(ROW0_THREADS)
score_row0_kernel(const Row0Params p) {
  extern __shared__ float sm[];
  float* x0 = sm;                                  // [2][NQ][4]
  float* h1 = x0 + 2 * p.NQ * ROW0_PAIRS;          // [2][128][4]
  float* part = h1 + 2 * HID * ROW0_PAIRS;         // [2 halves][2][128][4] partial pre-activations
  float* red = part + 2 * 2 * HID * ROW0_PAIRS;    // [2][4 warps][4]
  __shared__ int s_maxm;
  const int b0 = blockIdx.x * ROW0_PAIRS, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, H1n = p.NQ + 1;
  if (tid == 0) {
    int mm = 0;
    for (int i = 0; i < ROW0_PAIRS; ++i) if (b0 + i < p.B) mm = max(mm, p.matched_num[b0 + i]);
    s_maxm = mm;
  }
  for (int i = tid; i < 2 * p.NQ * ROW0_PAIRS; i += blockDim.x) x0[i] = 0.f;
  for (int i = tid; i < ROW0_PAIRS * p.NQp; i += blockDim.x) {      // geometry constants of every matched plane pair
    const int b = b0 + i / p.NQp, j = i % p.NQp;
    if (b < p.B) {
      float4 c0, c1, c2;
      const bool valid = j < p.matched_num[b];
      column_constants(p.geo_local + ((size_t)b * p.NQ + (valid ? j : 0)) * 6, valid, c0, c1, c2);
      float4* o = p.cjg + ((size_t)b * p.NQp + j) * 3;
      o[0] = c0; o[1] = c1; o[2] = c2;
    }
  }
  __syncthreads();
  const int maxm = s_maxm;
  {   // residual row of hypothesis 0: 4 warps per pair
    const int b = b0 + (warp >> 2), sub = warp & 3;
    if (b < p.B) {
      const int m = p.matched_num[b], pi = warp >> 2;
      const float* q = p.q0 + (size_t)b * 4;
      const float* t = p.t0 + (size_t)b * 3;
      const Mat3 Mr = quat_to_rot(q[0], q[1], q[2], q[3]);
      float R[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = Mr.m[i];
      const float* gl = p.geo_local + (size_t)b * p.NQ * 6;
      float sr = 0.f, st = 0.f;
      for (int j = sub * 32 + lane; j < m; j += 128) {
        float4 c0, c1, c2;
        column_constants(gl + (size_t)j * 6, true, c0, c1, c2);
        float xr, xt;
        residual_pair<true>(R, t[0], t[1], t[2], c0, c1, c2, xr, xt, sr, st);
        x0[(0 * p.NQ + j) * ROW0_PAIRS + pi] = xr;
        x0[(1 * p.NQ + j) * ROW0_PAIRS + pi] = xt;
      }
      sr = warp_sum(sr); st = warp_sum(st);
      if (lane == 0) { part[warp * 2] = sr; part[warp * 2 + 1] = st; }
    }
  }
  __syncthreads();
  if (tid < ROW0_PAIRS && b0 + tid < p.B) {      // masked distance sums of row 0 (fixed order over the 4 sub-warps)
    float a = 0.f, c = 0.f;
    for (int s4 = 0; s4 < 4; ++s4) { a += part[(tid * 4 + s4) * 2]; c += part[(tid * 4 + s4) * 2 + 1]; }
    p.sums[(size_t)(b0 + tid) * H1n] = a;
    p.sums[(size_t)p.B * H1n + (size_t)(b0 + tid) * H1n] = c;
  }
  __syncthreads();
  const int half = tid >> 8, br = (tid >> 7) & 1, t = tid & 127;
  float acc[ROW0_PAIRS];
  auto gemv = [&](const float* w, const float* x, int k0, int k1) {     // acc[i] += sum_k w[k][t] * x[k][i]
#pragma unroll
    for (int i = 0; i < ROW0_PAIRS; ++i) acc[i] = 0.f;
    int k = k0;
    for (; k + 8 <= k1; k += 8) {
      float wv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) wv[u] = __ldg(w + (size_t)(k + u) * HID);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float4 xa = *reinterpret_cast<const float4*>(x + (k + u) * ROW0_PAIRS);
        acc[0] = fmaf(wv[u], xa.x, acc[0]); acc[1] = fmaf(wv[u], xa.y, acc[1]); acc[2] = fmaf(wv[u], xa.z, acc[2]); acc[3] = fmaf(wv[u], xa.w, acc[3]);
      }
    }
    for (; k < k1; ++k) {
      const float wv = __ldg(w + (size_t)k * HID);
      const float4 xa = *reinterpret_cast<const float4*>(x + k * ROW0_PAIRS);
      acc[0] = fmaf(wv, xa.x, acc[0]); acc[1] = fmaf(wv, xa.y, acc[1]); acc[2] = fmaf(wv, xa.z, acc[2]); acc[3] = fmaf(wv, xa.w, acc[3]);
    }
  };
  {   // layer 1: the two halves of the K range run on different thread groups
    const int kmid = ((maxm + 1) / 2 + 7) / 8 * 8;
    const int k0 = half == 0 ? 0 : min(kmid, maxm), k1 = half == 0 ? min(kmid, maxm) : maxm;
    gemv(p.w1t + (size_t)br * p.NQ * HID + t, x0 + (size_t)br * p.NQ * ROW0_PAIRS, k0, k1);
#pragma unroll
    for (int i = 0; i < ROW0_PAIRS; ++i) part[((half * 2 + br) * HID + t) * ROW0_PAIRS + i] = acc[i];
  }
  __syncthreads();
  if (half == 0) {
    const float bias = p.vecs[br * HID + t];
#pragma unroll
    for (int i = 0; i < ROW0_PAIRS; ++i)
      h1[((size_t)br * HID + t) * ROW0_PAIRS + i] =
          fmaxf(part[((0 * 2 + br) * HID + t) * ROW0_PAIRS + i] + part[((1 * 2 + br) * HID + t) * ROW0_PAIRS + i] + bias, 0.f);
  }
  __syncthreads();
  {   // layer 2 (K = 128 split in halves) + folded layer 3 / regressor
    gemv(p.w2t + (size_t)br * HID * HID + t, h1 + (size_t)br * HID * ROW0_PAIRS, half * (HID / 2), (half + 1) * (HID / 2));
#pragma unroll
    for (int i = 0; i < ROW0_PAIRS; ++i) part[((half * 2 + br) * HID + t) * ROW0_PAIRS + i] = acc[i];
  }
  __syncthreads();
  if (half == 0) {
    const float bias = p.vecs[2 * HID + br * HID + t], w34 = p.vecs[4 * HID + br * HID + t];
#pragma unroll
    for (int i = 0; i < ROW0_PAIRS; ++i) {
      const float pre = part[((0 * 2 + br) * HID + t) * ROW0_PAIRS + i] + part[((1 * 2 + br) * HID + t) * ROW0_PAIRS + i] + bias;
      const float v = warp_sum(fmaxf(pre, 0.f) * w34);
      if (lane == 0) red[(br * 4 + (warp & 3)) * ROW0_PAIRS + i] = v;
    }
  }
  __syncthreads();
  if (tid < 2 * ROW0_PAIRS) {
    const int br2 = tid / ROW0_PAIRS, i = tid % ROW0_PAIRS, b = b0 + i;
    if (b < p.B) {
      float v = 0.f;
      for (int w4 = 0; w4 < 4; ++w4) v += red[(br2 * 4 + w4) * ROW0_PAIRS + i];
      p.logits[(size_t)br2 * p.B * H1n + (size_t)b * H1n] = v;     // like the tile logits: without the constant c34
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
    redi[0] = bi;
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
  if (tid < 2) misc[tid] = p.logits[(size_t)tid * p.B * H1n + (size_t)b * H1n];     // hypothesis 0 (score_row0_kernel)
  __syncthreads();
  // ---- merge the tile partials with hypothesis 0 (log-sum-exp rescale)
  const int ntile = (m + TILE_H - 1) / TILE_H;
  if (tid < 2) {
    const int br = tid;
    float M = misc[br];                      // every logit here lacks the constant c34 (softmax-invariant)
    for (int t = 0; t < ntile; ++t) M = fmaxf(M, p.partials[(((size_t)b * p.tiles_per_pair + t) * 2 + br) * PART_STRIDE]);
    float S = expf(misc[br] - M);
    for (int t = 0; t < ntile; ++t) {
      const float* part = p.partials + (((size_t)b * p.tiles_per_pair + t) * 2 + br) * PART_STRIDE;
      S += expf(part[0] - M) * part[1];
    }
    misc[2 + br] = M;
    misc[4 + br] = S;
  }
  __syncthreads();
  // scores (softmax over hypotheses 0..m, :1010-1014, :1039-1043) -> global (optional) + kept for max-score
  const float Mr_ = misc[2], Mt_ = misc[3], Sr_ = misc[4], St_ = misc[5];
  float* lr = p.logits + (size_t)b * H1n;
  float* lt = p.logits + (size_t)p.B * H1n + (size_t)b * H1n;
  for (int h = tid; h <= m; h += blockDim.x) {
    const float a = expf(lr[h] - Mr_) / Sr_;
    const float c = expf(lt[h] - Mt_) / St_;
    lr[h] = a; lt[h] = c;                     // logits buffer now holds the scores
    if (p.score_rot) p.score_rot[(size_t)b * H1n + h] = a;
    if (p.score_tran) p.score_tran[(size_t)b * H1n + h] = c;
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
  // ---- aggregated features, one channel per thread (:1047-1087)
  {
    const int c = tid;
    const float f0r = p.feat_rot0[(size_t)b * C + c], f0t = p.feat_tran0[(size_t)b * C + c];
    float ar = 0.f, at = 0.f, wr = 0.f, wt = 0.f;
    for (int t = 0; t < ntile; ++t) {
      const float* pr = p.partials + (((size_t)b * p.tiles_per_pair + t) * 2 + 0) * PART_STRIDE;
      const float* pt = pr + PART_STRIDE;
      wr = fmaf(expf(pr[0] - Mr_), pr[PART_HDR + c], wr);
      wt = fmaf(expf(pt[0] - Mt_), pt[PART_HDR + c], wt);
      ar += pr[PART_HDR + C + c];
      at += pt[PART_HDR + C + c];
    }
    wr = (wr + expf(misc[0] - Mr_) * f0r) / Sr_;
    wt = (wt + expf(misc[1] - Mt_) * f0t) / St_;
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

extern "C" size_t nsac_score_pack_bytes(int NQ) {
  return NQ < 1 ? 0 : align256(pack_total_bytes(NQ, nq_padded(NQ)));
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
         (size_t)B * nq_padded(NQ) * 48 + 256;
}

extern "C" int nsac_score_aggregate_tc(const float* geo_local, const float* q_h, const float* t_h, const float* q0,
                                       const float* t0, const float* feat_rot, const float* feat_tran,
                                       const float* feat_rot0, const float* feat_tran0, const int32_t* matched_num,
                                       const void* pack, const float* w_rots, const float* b_rots, const float* w_trans,
                                       const float* b_trans, int B, int NQ, int out_cam_type, float* pose, float* score_rot,
                                       float* score_tran, int32_t* sel_idx, void* workspace,
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
                   (reinterpret_cast<uintptr_t>(q_h) & 15) == 0,
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
  float4* cjg = reinterpret_cast<float4*>(ws + align256(4 * per * sizeof(float)) + align256((size_t)B * tiles * 2 * PART_STRIDE * sizeof(float)));
  const float* vecs = reinterpret_cast<const float*>(pk + pack_off_vecs(NQp));

  // hypothesis 0 of every pair (independent of the tile kernel)
  Row0Params rp;
  rp.geo_local = geo_local; rp.q0 = q0; rp.t0 = t0; rp.matched_num = matched_num;
  rp.w1t = reinterpret_cast<const float*>(pk + pack_off_w1t(NQp));
  rp.w2t = reinterpret_cast<const float*>(pk + pack_off_w2t(NQ, NQp));
  rp.vecs = vecs; rp.B = B; rp.NQ = NQ; rp.NQp = NQp; rp.logits = logits; rp.sums = sums; rp.cjg = cjg;
  const size_t row0_smem = sizeof(float) * ((size_t)2 * NQ * ROW0_PAIRS + 2 * HID * ROW0_PAIRS + 4 * HID * ROW0_PAIRS + 2 * 4 * ROW0_PAIRS);
  NSAC_REQUIRE(row0_smem <= 200 * 1024, "nsac_score_aggregate_tc: NQ=%d too large", NQ);
  if (row0_smem > 48 * 1024)
    NSAC_CUDA(cudaFuncSetAttribute(score_row0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row0_smem));
  score_row0_kernel<<<nsac_cdiv(B, ROW0_PAIRS), ROW0_THREADS, row0_smem, s>>>(rp);
  NSAC_CHECK_LAUNCH("score_row0_kernel");

  CUtensorMap m1r, m1t, m2r, m2t;
  const bool ok = make_map_f16(&m1r, pk, HID, NQp) && make_map_f16(&m1t, pk + pack_w1_bytes(NQp), HID, NQp) &&
                  make_map_f16(&m2r, pk + 2 * pack_w1_bytes(NQp), HID, HID) &&
                  make_map_f16(&m2t, pk + 2 * pack_w1_bytes(NQp) + HID * HID * 2, HID, HID);
  if (!ok) {
    nsac_set_error("nsac_score_aggregate_tc: cuTensorMapEncodeTiled failed");
    return NSAC_ERR_LAUNCH;
  }
  TcParams tp;
  tp.geo_local = geo_local; tp.q_h = q_h; tp.t_h = t_h; tp.feat_rot = feat_rot; tp.feat_tran = feat_tran;
  tp.matched_num = matched_num; tp.vecs = vecs; tp.cjg = cjg; tp.B = B; tp.NQ = NQ; tp.NQp = NQp; tp.tiles_per_pair = tiles;
  tp.need_sums = out_cam_type == NSAC_CAM_MIN_COST; tp.logits = logits; tp.sums = sums; tp.partials = partials;
  static bool attr = false;
  if (!attr) {
    NSAC_CUDA(cudaFuncSetAttribute(score_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    NSAC_CUDA(cudaFuncSetAttribute(score_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr = true;
  }
  const int items = B * tiles;
  const int grid = items < sm_count() ? items : sm_count();
  if (tp.need_sums)
    score_tc_kernel<true><<<grid, NUM_THREADS, SMEM_BYTES, s>>>(m1r, m1t, m2r, m2t, tp);
  else
    score_tc_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, s>>>(m1r, m1t, m2r, m2t, tp);
  NSAC_CHECK_LAUNCH("score_tc_kernel");

  SelParams sp;
  sp.q_h = q_h; sp.t_h = t_h; sp.q0 = q0; sp.t0 = t0; sp.feat_rot0 = feat_rot0; sp.feat_tran0 = feat_tran0;
  sp.matched_num = matched_num; sp.w_rots = w_rots; sp.b_rots = b_rots; sp.w_trans = w_trans; sp.b_trans = b_trans;
  sp.logits = logits; sp.sums = sums; sp.partials = partials;
  sp.B = B; sp.NQ = NQ; sp.tiles_per_pair = tiles; sp.out_cam_type = out_cam_type; sp.pose = pose; sp.score_rot = score_rot;
  sp.score_tran = score_tran; sp.sel_idx = sel_idx;
  sp.peer_rows = peer_rows; sp.num_peers = num_peers; sp.row_offset = row_offset;
  const size_t sel_smem = sizeof(float) * (4 * C_FEAT + SEL_THREADS + 16 + 8) + sizeof(int) * SEL_THREADS;
  score_select_tc_kernel<<<B, SEL_THREADS, sel_smem, s>>>(sp);
  NSAC_CHECK_LAUNCH("score_select_tc_kernel");
  return NSAC_OK;
}
